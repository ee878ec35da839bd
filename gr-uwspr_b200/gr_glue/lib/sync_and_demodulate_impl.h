// uwspr.sync_and_demodulate on the GPU: refinement chain and soft symbols from uwspr_b200_fine(),
// de-interleaver + Fano decoder on the host (uwspr_b200_decode_candidate), one blob PDU per decoded
// candidate, and the reference's messagelog.txt.
#ifndef INCLUDED_UWSPR_SYNC_AND_DEMODULATE_IMPL_H
#define INCLUDED_UWSPR_SYNC_AND_DEMODULATE_IMPL_H

#include <uwspr/sync_and_demodulate.h>

#include <stdio.h>
#include <time.h>

#include <vector>

#include "uwspr_b200.h"

namespace gr {
namespace uwspr {

class sync_and_demodulate_impl : public sync_and_demodulate
{
public:
    sync_and_demodulate_impl(int fs, int fl, int spb, int maxdrift, int maxfreqs, int cf);
    ~sync_and_demodulate_impl();
    // handler of message port "in" (reference: sync_and_demodulate_impl::demodulate, lib/sync_and_demodulate_impl.cc:315-534)
    void demodulate(pmt::pmt_t msg);
    int framecount() const { return d_framecount; }

private:
    void log_frame(const uwspr_b200_candidate_t &cand, const int8_t message7[7]);
    pmt::pmt_t d_in_port, d_out_port;
    uwspr_b200_ctx *d_ctx;
    int d_fl, d_maxfreqs, d_framecount;
    float *d_iq;   // pinned, 2*fl floats
    std::vector<uwspr_b200_candidate_t> d_cands, d_retry;
    std::vector<uwspr_b200_refined_t> d_refined;
    std::vector<uwspr_b200_jiggle_t> d_jig;
    std::vector<uint8_t> d_soft;
    FILE *d_log;
    time_t d_start;
};

}  // namespace uwspr
}  // namespace gr

#endif
