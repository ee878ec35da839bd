/* Oracle stub for <gnuradio/sync_block.h> (test infrastructure). */
#ifndef ORACLE_STUB_GR_SYNC_BLOCK_H
#define ORACLE_STUB_GR_SYNC_BLOCK_H
#include <gnuradio/block.h>
namespace gr {
class sync_block : virtual public block
{
public:
    sync_block() {}
    sync_block(const std::string &name, io_signature::sptr i, io_signature::sptr o) : block(name, i, o) {}
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) = 0;
};
} // namespace gr
#endif
