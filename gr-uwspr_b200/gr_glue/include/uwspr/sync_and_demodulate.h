/* uwspr.sync_and_demodulate -- public class of the fine-sync / demodulation / decoding block
 * (reference: include/uwspr/sync_and_demodulate.h:33-49, grc/uwspr_sync_and_demodulate.xml:7). */
#ifndef INCLUDED_UWSPR_SYNC_AND_DEMODULATE_H
#define INCLUDED_UWSPR_SYNC_AND_DEMODULATE_H

#include <gnuradio/block.h>
#include <uwspr/api.h>

namespace gr {
namespace uwspr {

class UWSPR_API sync_and_demodulate : virtual public gr::block
{
public:
    typedef boost::shared_ptr<sync_and_demodulate> sptr;
    static sptr make(int fs, int fl, int spb, int maxdrift, int maxfreqs, int cf);
};

}  // namespace uwspr
}  // namespace gr

#endif
