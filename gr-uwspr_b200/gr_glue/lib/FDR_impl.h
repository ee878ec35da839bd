// uwspr.FDR on the GPU: the block's message handler marshals the window PDU into one
// uwspr_b200_coarse() call (libuwspr_b200.so) and publishes the candidate PDU the reference publishes.
#ifndef INCLUDED_UWSPR_FDR_IMPL_H
#define INCLUDED_UWSPR_FDR_IMPL_H

#include <uwspr/FDR.h>

#include "uwspr_b200.h"

namespace gr {
namespace uwspr {

class FDR_impl : public FDR
{
public:
    FDR_impl(int fs, int fl, int spb, int maxdrift, int maxfreqs, int halfbandwidth, int cf, int threshold);
    ~FDR_impl();
    // handler of message port "in" (reference: FDR_impl::transform, lib/FDR_impl.cc:214-456)
    void transform(pmt::pmt_t msg);

private:
    pmt::pmt_t d_in_port, d_out_port;
    uwspr_b200_ctx *d_ctx;
    int d_fl, d_maxfreqs;
    float *d_iq;                      // pinned, 2*fl floats
    uwspr_b200_candidate_t *d_cands;  // pinned, maxfreqs records
};

}  // namespace uwspr
}  // namespace gr

#endif
