/* uwspr_oracle -- CPU restatement of the gr-uwspr receive hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (gr-uwspr_b200/) never includes, links or calls anything under oracle/.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the
 * reference's own object code (oracle/_ref, the unmodified reference sources built
 * against stubs) on the reference fixtures and on seeded synthetic windows by
 * tests/test_oracle_vs_reference.py (run where /root/reference exists) and against
 * the committed golden vectors in tests/golden/ everywhere else.  The one piece
 * that cannot be pinned is the FFT itself: the reference calls FFTW3f (unpinned
 * third-party, absent here); orc_spectrogram() evaluates the DFT definition in
 * double precision and rounds once to float, as oracle/stubs/fftw3_stub.c does.
 *
 * All citations are file:line under /root/reference.
 */
#ifndef UWSPR_ORACLE_H
#define UWSPR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/candidate_t.h:27-50 -- 48 bytes, alignment 8 */
typedef struct {
    float freq;
    float snr;
    float drift; /* never written by the reference */
    float sync;
    int32_t shift;
    int32_t m_type; /* 0 = linear, 1 = nonlinear */
    union {
        struct { float drift; } lin;
        struct { double V1, V2; int32_t p1, p2; } nl;
    } u;
} orc_candidate_t;

/* Derived constants of FDR_impl::FDR_impl (lib/FDR_impl.cc:48-151) */
typedef struct {
    int fs, fl, spb, maxdrift, maxfreqs, halfbandwidth, cf;
    float threshold;
    int size;  /* DFT length, 2*spb                  :81  */
    int m;     /* index of the DC bin after the shift :95  */
    int hpbm;  /* pass band half width in bins        :97  */
    int n;     /* number of half-symbol rows          :109 */
    float df;  /* bin spacing                         :93  */
    float min_snr; /*                                 :137 */
    float w[4096]; /* half-sine window                :103-105 (size <= 4096) */
} orc_fdr_t;

/* returns 0, or -1 where the reference would exit() (:85-90) or index out of
 * bounds (hazard H3: hpbm+3 > m) */
int orc_fdr_init(orc_fdr_t *f, int fs, int fl, int spb, int maxdrift, int maxfreqs,
                 int halfbandwidth, int cf, int threshold);

/* lib/FDR_impl.cc:222-254.  iq: fl interleaved complex64.  ps: [n][size].
 * spectra_out (may be NULL): [n][size] complex64 in FFT order (bin 0 = DC). */
void orc_spectrogram(const orc_fdr_t *f, const float *iq, float *ps, float *spectra_out);
/* :246-253 only: ps from given spectra [n][size] complex64 (FFT order) */
void orc_power(const orc_fdr_t *f, const float *spectra, float *ps);

/* lib/FDR_impl.cc:257-319.  psavg [size], smspec [2*hpbm] may be NULL.
 * cands: maxfreqs records; returns npk. */
int orc_normalize_peaks(const orc_fdr_t *f, const float *ps, float *psavg, float *smspec,
                        orc_candidate_t *cands);

/* lib/FDR_impl.cc:339-409 (+ powersum :188-210) */
void orc_coarse(const orc_fdr_t *f, const float *ps, orc_candidate_t *cands, int npk);

/* whole FDR_impl::transform; returns npk */
int orc_fdr_transform(const orc_fdr_t *f, const float *iq, orc_candidate_t *cands, float *ps_scratch);

/* lib/slm.cc:36-73 */
float orc_slm_frequency_drift(double V1, double V2, int p1, int p2, float cf, float t);
/* lib/slm.cc:76-116: k-th trajectory of the generator, k in [0,125); returns 0 past the end */
int orc_slm_trajectory(int k, double *V1, double *V2, int *p1, int *p2);

/* lib/sync_and_demodulate_impl.cc:126-256.  The uninitialised `t` of the nonlinear branch (:177-180, hazard H1); the compiled
 * reference behaves as t == 0 (orc_set_nonlinear_intended_t(1) selects i*111/162). */
void orc_set_nonlinear_intended_t(int on);
void orc_sync_and_demodulate(const orc_candidate_t *cand, int cf, const float *id, const float *qd,
                             long np, unsigned char *symbols, float *f1, int ifmin, int ifmax,
                             float fstep, int *shift1, int lagmin, int lagmax, int lagstep,
                             float *drift1, int symfac, float *sync, int mode);

/* one refinement call as seen by the driver (same fields as the harness trace) */
typedef struct {
    int mode, lagmin, lagmax, lagstep, ifmin, ifmax;
    float fstep;
    float f1_in;
    int shift_in;
    float drift_in;
    float f1_out;
    int shift_out;
    float sync_out;
    unsigned char symbols[162];
    unsigned char pad[2];
} orc_sd_call_t;

typedef struct {
    unsigned char symbols[162]; /* deinterleaved, as handed to the decoder */
    unsigned char data[11];
    unsigned char pad[3];
    int result;
    unsigned int metric, cycles, maxnp;
} orc_fano_call_t;

typedef struct {
    orc_sd_call_t *calls;
    int max_calls, n_calls;
    orc_fano_call_t *fanos;
    int max_fanos, n_fanos;
} orc_trace_t;

/* lib/sync_and_demodulate_impl.cc:389-531 for the npk candidates of one window.
 * blobs: 7 bytes per decoded candidate; returns the number of blobs.
 * run_fano == 0 skips the decoder (treated as "not decoded"), which makes the
 * driver evaluate all 17 jiggled shifts of every gated candidate. */
int orc_demodulate(int cf, const float *iq, int fl, const orc_candidate_t *cands, int npk,
                   orc_trace_t *trace, unsigned char *blobs, int max_blobs, int run_fano);

/* same, also reporting per candidate f1, shift1, drift1, sync1, worth_a_try as they stand
 * when the peak-up loop starts (:457); refined: [npk][5] floats or NULL */
int orc_demodulate_ex(int cf, const float *iq, int fl, const orc_candidate_t *cands, int npk,
                      orc_trace_t *trace, unsigned char *blobs, int max_blobs, int run_fano,
                      float *refined);

/* lib/sync_and_demodulate_impl.cc:265-282 */
void orc_deinterleave(unsigned char *sym162);
/* inverse permutation (transmit side) */
void orc_interleave(unsigned char *sym162);

/* lib/Fano.cc:81-100 */
void orc_encode(unsigned char *symbols, const unsigned char *data, unsigned int nbytes);
/* lib/Fano.cc:110-252; returns 0 on success, -1 on time-out */
int orc_fano(unsigned int *metric, unsigned int *cycles, unsigned int *maxnp, unsigned char *data,
             const unsigned char *symbols, unsigned int nbits, int delta, unsigned int maxcycles);

/* 162 channel symbols (0..3) of a 7-byte packed message (+4 zero tail bytes):
 * encode -> interleave -> 2*data + sync (verified against examples/VE3EMB.c2) */
void orc_channel_symbols(const unsigned char *msg7, unsigned char *chan162);

int orc_sync_bit(int i);

/* lib/sliding_window_stream_to_pdu_impl.cc:98-138: window k = stream[k*shift*fs, +fl).
 * state-free restatement: number of windows emitted after `nitems` items arrive in
 * calls of `chunk` items (one PDU at most per work() call). */
long orc_sliding_window_count(long nitems, int chunk, int fs, int fl, int shift);

#ifdef __cplusplus
}
#endif
#endif
