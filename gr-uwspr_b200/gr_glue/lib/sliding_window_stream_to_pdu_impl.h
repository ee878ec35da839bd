// uwspr.sliding_window_stream_to_pdu: the stream side of the receive chain.  Window k of the stream is
// stream[k*shift*fs, k*shift*fs + fl); at most one window PDU is published per work() call, as in the
// reference (lib/sliding_window_stream_to_pdu_impl.cc:98-138).  Host-only: no device work happens here.
#ifndef INCLUDED_UWSPR_SLIDING_WINDOW_STREAM_TO_PDU_IMPL_H
#define INCLUDED_UWSPR_SLIDING_WINDOW_STREAM_TO_PDU_IMPL_H

#include <uwspr/sliding_window_stream_to_pdu.h>

#include <vector>

namespace gr {
namespace uwspr {

class sliding_window_stream_to_pdu_impl : public sliding_window_stream_to_pdu
{
public:
    sliding_window_stream_to_pdu_impl(int fs, int fl, int shift, int C);
    ~sliding_window_stream_to_pdu_impl();
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);

private:
    pmt::pmt_t d_out_port;
    int d_fl, d_step;
    size_t d_capacity;            // ring capacity in samples (C windows)
    std::vector<gr_complex> d_ring;
    size_t d_head, d_size;        // oldest sample, samples held
    long d_ready;                 // samples not yet consumed by a published window
};

}  // namespace uwspr
}  // namespace gr

#endif
