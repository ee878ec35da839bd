// PMT <-> plain-data conversion of the PDUs the uwspr blocks exchange:
//   (1) window      cons(PMT_NIL, vector[fl] of pmt complex)            sliding window -> FDR
//   (2) candidates  cons(PMT_NIL, tuple(window vector, long npk, vector[npk] of tuple))   FDR -> sync_and_demodulate
//         linear    tuple(long 0, double freq, double snr, double sync, long shift, double drift)
//         nonlinear tuple(long 1, double freq, double snr, double sync, long shift, double V1, double V2, long p1, long p2)
//   (3) message     cons(PMT_NIL, blob[7])                                sync_and_demodulate -> WSPR_unpacker
// as the reference writes them (lib/FDR_impl.cc:414-455, lib/sliding_window_stream_to_pdu_impl.cc:106-131,
// lib/sync_and_demodulate_impl.cc:528-530) and reads them back (lib/FDR_impl.cc:218-231,
// lib/sync_and_demodulate_impl.cc:337-377).
#ifndef UWSPR_B200_GR_GLUE_PDU_CODEC_H
#define UWSPR_B200_GR_GLUE_PDU_CODEC_H

#include <pmt/pmt.h>

#include <stdexcept>
#include <string>

#include "uwspr_b200.h"

namespace gr {
namespace uwspr {
namespace glue {

// the fl samples of a window vector as interleaved fp32 (I, Q); the PMTs hold fp32 values widened to double
inline void window_from_pmt(const pmt::pmt_t &vec, int fl, float *iq)
{
    for (int i = 0; i < fl; i++) {
        const std::complex<double> c = pmt::to_complex(pmt::vector_ref(vec, i));
        iq[2 * i] = (float)c.real();
        iq[2 * i + 1] = (float)c.imag();
    }
}

inline pmt::pmt_t candidate_to_pmt(const uwspr_b200_candidate_t &c)
{
    if (c.m_type == 0)
        return pmt::make_tuple(pmt::from_long(0), pmt::from_double(c.freq), pmt::from_double(c.snr), pmt::from_double(c.sync),
                               pmt::from_long(c.shift), pmt::from_double(c.m_linear.drift));
    return pmt::make_tuple(pmt::from_long(1), pmt::from_double(c.freq), pmt::from_double(c.snr), pmt::from_double(c.sync),
                           pmt::from_long(c.shift), pmt::from_double(c.m_nonlinear.V1), pmt::from_double(c.m_nonlinear.V2),
                           pmt::from_long(c.m_nonlinear.p1), pmt::from_long(c.m_nonlinear.p2));
}

inline pmt::pmt_t candidates_pdu(const pmt::pmt_t &window_vec, const uwspr_b200_candidate_t *cands, int npk)
{
    pmt::pmt_t list = pmt::make_vector(npk, pmt::PMT_NIL);
    for (int i = 0; i < npk; i++) pmt::vector_set(list, i, candidate_to_pmt(cands[i]));
    // the window vector is forwarded as the same object, not copied
    return pmt::cons(pmt::PMT_NIL, pmt::make_tuple(window_vec, pmt::from_long(npk), list));
}

inline uwspr_b200_candidate_t candidate_from_pmt(const pmt::pmt_t &t)
{
    uwspr_b200_candidate_t c;
    memset(&c, 0, sizeof(c));
    c.m_type = (int32_t)pmt::to_long(pmt::tuple_ref(t, 0));
    c.freq = (float)pmt::to_double(pmt::tuple_ref(t, 1));
    c.snr = (float)pmt::to_double(pmt::tuple_ref(t, 2));
    c.sync = (float)pmt::to_double(pmt::tuple_ref(t, 3));
    c.shift = (int32_t)pmt::to_long(pmt::tuple_ref(t, 4));
    if (c.m_type == 0) {
        c.m_linear.drift = (float)pmt::to_double(pmt::tuple_ref(t, 5));
    } else if (c.m_type == 1) {
        c.m_nonlinear.V1 = pmt::to_double(pmt::tuple_ref(t, 5));
        c.m_nonlinear.V2 = pmt::to_double(pmt::tuple_ref(t, 6));
        c.m_nonlinear.p1 = (int32_t)pmt::to_long(pmt::tuple_ref(t, 7));
        c.m_nonlinear.p2 = (int32_t)pmt::to_long(pmt::tuple_ref(t, 8));
        // the reference then stores 0 into m_linear.drift, which shares storage with the low half of V1
        // (sync_and_demodulate_impl.cc:373); the library applies the same overwrite inside the fine stage
    } else {
        throw std::runtime_error("uwspr: candidate tuple with unknown drift model " + std::to_string(c.m_type));
    }
    return c;
}

inline pmt::pmt_t message_pdu(const int8_t message7[7]) { return pmt::cons(pmt::PMT_NIL, pmt::make_blob(message7, 7)); }

}  // namespace glue
}  // namespace uwspr
}  // namespace gr

#endif
