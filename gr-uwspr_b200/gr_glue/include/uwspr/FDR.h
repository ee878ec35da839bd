/* uwspr.FDR -- public class of the frequency-domain receiver block.
 *
 * Same class, base and factory signature as the reference (include/uwspr/FDR.h:33-50,
 * grc/uwspr_FDR.xml:7), so flowgraphs and the SWIG wrapper (swig/uwspr_swig.i:20-29) bind
 * unchanged; the implementation behind make() runs on the GPU through libuwspr_b200.so. */
#ifndef INCLUDED_UWSPR_FDR_H
#define INCLUDED_UWSPR_FDR_H

#include <gnuradio/block.h>
#include <uwspr/api.h>

namespace gr {
namespace uwspr {

class UWSPR_API FDR : virtual public gr::block
{
public:
    typedef boost::shared_ptr<FDR> sptr;
    /* fs: sample rate (375), fl: window length in samples (45000), spb: samples per symbol (256),
     * maxdrift: largest linear drift searched (Hz), maxfreqs: candidate cap, halfbandwidth: half pass
     * band (Hz), cf: carrier (Hz) of the straight-line Doppler model, threshold: ratio a straight-line
     * hypothesis must beat the running best by */
    static sptr make(int fs, int fl, int spb, int maxdrift, int maxfreqs, int halfbandwidth, int cf, int threshold);
};

}  // namespace uwspr
}  // namespace gr

#endif
