/* uwspr.sliding_window_stream_to_pdu -- stream in, one PDU of fl samples out every `shift` seconds
 * (reference: include/uwspr/sliding_window_stream_to_pdu.h:37-53, grc/uwspr_sliding_window_stream_to_pdu.xml:7). */
#ifndef INCLUDED_UWSPR_SLIDING_WINDOW_STREAM_TO_PDU_H
#define INCLUDED_UWSPR_SLIDING_WINDOW_STREAM_TO_PDU_H

#include <gnuradio/sync_block.h>
#include <uwspr/api.h>

namespace gr {
namespace uwspr {

class UWSPR_API sliding_window_stream_to_pdu : virtual public gr::sync_block
{
public:
    typedef boost::shared_ptr<sliding_window_stream_to_pdu> sptr;
    /* fs: sample rate, fl: window length (samples), shift: seconds between windows, C: ring capacity in windows */
    static sptr make(int fs, int fl, int shift, int C);
};

}  // namespace uwspr
}  // namespace gr

#endif
