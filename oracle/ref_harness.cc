/* Oracle harness around the UNMODIFIED reference sources (test infrastructure,
 * not product code; nothing under gr-uwspr_b200/ may link or call this).
 *
 * The reference translation units (lib/FDR_impl.cc, lib/sync_and_demodulate_impl.cc,
 * lib/slm.cc, lib/Fano.cc, lib/tab.c, lib/sliding_window_stream_to_pdu_impl.cc) are
 * compiled where they lie under /root/reference into oracle/_ref/libuwspr_ref.so
 * against the stub headers in oracle/stubs/.  This file is linked into a second
 * shared object, oracle/_ref/libref_harness.so, which
 *   - drives the blocks' message handlers synchronously with PDUs built from
 *     plain arrays (schemas: FDR_impl.cc:218-221,414-455;
 *     sync_and_demodulate_impl.cc:337-377,528-530),
 *   - reads block internals (ps, psavg, candidates) by including the two
 *     *_impl.h headers with `private` spelled `public` (layout is unchanged),
 *   - interposes, by ELF symbol interposition, the two functions whose
 *     arguments are otherwise unobservable: sync_and_demodulate_impl::
 *     sync_and_demodulate (every refinement call of the driver) and Fano::fano
 *     (soft symbols as handed to the decoder, decoder outcome and time).
 *     The originals are reached with dlsym(RTLD_NEXT).
 */
#define _GNU_SOURCE 1
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <complex>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>

#include <gnuradio/block.h>
#include <gnuradio/sync_block.h>
#include <fftw3.h>

#define private public
#include "FDR_impl.h"
#include "sync_and_demodulate_impl.h"
#include "sliding_window_stream_to_pdu_impl.h"
#include "helpers.h"
#undef private

using namespace gr::uwspr;

extern "C" {

typedef struct {
    int mode, lagmin, lagmax, lagstep, ifmin, ifmax;
    float fstep;
    float f1_in;
    int shift_in;
    float drift_in;
    float f1_out;
    int shift_out;
    float sync_out;
    unsigned char symbols[162]; /* mode 2: soft symbols as returned (before deinterleave) */
    unsigned char pad[2];
} ref_sd_call_t;

typedef struct {
    unsigned char symbols[162]; /* deinterleaved soft symbols handed to the decoder */
    unsigned char data[11];
    unsigned char pad[3];
    int result;
    unsigned int metric, cycles, maxnp;
} ref_fano_call_t;

typedef struct {
    ref_sd_call_t *calls;
    int max_calls, n_calls;
    ref_fano_call_t *fanos;
    int max_fanos, n_fanos;
} ref_trace_t;

} // extern "C"

static ref_trace_t *g_trace = NULL;
static double g_fano_seconds = 0.0;
static long g_fano_calls = 0;
static int g_skip_fano = 0; /* 1: do not run the decoder, report "not decoded" */

static double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---- interposers ------------------------------------------------------- */
namespace gr {
namespace uwspr {

typedef void (*real_sd_t)(sync_and_demodulate_impl *, candidate_t, float *, float *, long,
                          unsigned char *, float *, int, int, float, int *, int, int, int, float *,
                          int, float *, int);

void sync_and_demodulate_impl::sync_and_demodulate(candidate_t candidate, float *id, float *qd,
                                                   long np, unsigned char *symbols, float *f1,
                                                   int ifmin, int ifmax, float fstep, int *shift1,
                                                   int lagmin, int lagmax, int lagstep,
                                                   float *drift1, int symfac, float *sync, int mode)
{
    static real_sd_t real = NULL;
    if (!real) {
        real = (real_sd_t)dlsym(
            RTLD_NEXT,
            "_ZN2gr5uwspr24sync_and_demodulate_impl19sync_and_demodulateENS0_11candidate_tEPfS3_lPhS3_iifPiiiiS3_iS3_i");
        if (!real) {
            fprintf(stderr, "ref_harness: cannot resolve reference sync_and_demodulate: %s\n", dlerror());
            abort();
        }
    }
    ref_sd_call_t rec;
    memset(&rec, 0, sizeof(rec));
    rec.mode = mode;
    rec.lagmin = lagmin;
    rec.lagmax = lagmax;
    rec.lagstep = lagstep;
    rec.ifmin = ifmin;
    rec.ifmax = ifmax;
    rec.fstep = fstep;
    rec.f1_in = *f1;
    rec.shift_in = *shift1;
    rec.drift_in = *drift1;
    real(this, candidate, id, qd, np, symbols, f1, ifmin, ifmax, fstep, shift1, lagmin, lagmax,
         lagstep, drift1, symfac, sync, mode);
    rec.f1_out = *f1;
    rec.shift_out = *shift1;
    rec.sync_out = *sync;
    if (mode == 2) memcpy(rec.symbols, symbols, 162);
    if (g_trace && g_trace->calls && g_trace->n_calls < g_trace->max_calls)
        g_trace->calls[g_trace->n_calls] = rec;
    if (g_trace) g_trace->n_calls++;
}

typedef int (*real_fano_t)(Fano *, unsigned int *, unsigned int *, unsigned int *, unsigned char *,
                           unsigned char *, unsigned int, int (*)[256], int, unsigned int);

int Fano::fano(unsigned int *metric, unsigned int *cycles, unsigned int *maxnp, unsigned char *data,
               unsigned char *symbols, unsigned int nbits, int mettab[2][256], int delta,
               unsigned int maxcycles)
{
    static real_fano_t real = NULL;
    if (!real) {
        real = (real_fano_t)dlsym(RTLD_NEXT, "_ZN2gr5uwspr4Fano4fanoEPjS2_S2_PhS3_jPA256_iij");
        if (!real) {
            fprintf(stderr, "ref_harness: cannot resolve reference Fano::fano: %s\n", dlerror());
            abort();
        }
    }
    ref_fano_call_t rec;
    memset(&rec, 0, sizeof(rec));
    memcpy(rec.symbols, symbols, 162);
    int r;
    double t0 = now_s();
    if (g_skip_fano) {
        *metric = 0;
        *cycles = 0;
        *maxnp = 0;
        r = -1;
    } else {
        r = real(this, metric, cycles, maxnp, data, symbols, nbits, mettab, delta, maxcycles);
    }
    g_fano_seconds += now_s() - t0;
    g_fano_calls++;
    rec.result = r;
    rec.metric = *metric;
    rec.cycles = *cycles;
    rec.maxnp = *maxnp;
    if (!g_skip_fano) memcpy(rec.data, data, 11);
    if (g_trace && g_trace->fanos && g_trace->n_fanos < g_trace->max_fanos)
        g_trace->fanos[g_trace->n_fanos] = rec;
    if (g_trace) g_trace->n_fanos++;
    return r;
}

} // namespace uwspr
} // namespace gr

/* ---- FFT injection ----------------------------------------------------- */
struct inject_t {
    const float *spectra; /* [rows][n][2], FFT natural order (DC first) */
    int rows, row;
};
static int inject_hook(void *user, int n, const fftwf_complex *in, fftwf_complex *out)
{
    (void)in;
    inject_t *s = (inject_t *)user;
    if (!s->spectra || s->row >= s->rows) return 0;
    memcpy(out, s->spectra + (size_t)s->row * n * 2, sizeof(float) * 2 * n);
    s->row++;
    return 1;
}

/* ---- helpers ----------------------------------------------------------- */
static pmt::pmt_t samples_to_pmt(const float *iq, int fl)
{
    pmt::pmt_t v = pmt::make_vector(fl, pmt::PMT_NIL);
    for (int i = 0; i < fl; i++)
        pmt::vector_set(v, i, pmt::make_rectangular(iq[2 * i], iq[2 * i + 1]));
    return v;
}

static pmt::pmt_t candidates_to_pmt(const candidate_t *c, int npk)
{
    /* FDR_impl.cc:416-447 */
    pmt::pmt_t v = pmt::make_vector(npk, pmt::PMT_NIL);
    for (int i = 0; i < npk; i++) {
        pmt::pmt_t t;
        if (c[i].m_type == linear) {
            t = pmt::make_tuple(pmt::from_long(linear), pmt::from_double(c[i].freq),
                                pmt::from_double(c[i].snr), pmt::from_double(c[i].sync),
                                pmt::from_long(c[i].shift), pmt::from_double(c[i].m_linear.drift));
        } else {
            t = pmt::make_tuple(pmt::from_long(nonlinear), pmt::from_double(c[i].freq),
                                pmt::from_double(c[i].snr), pmt::from_double(c[i].sync),
                                pmt::from_long(c[i].shift), pmt::from_double(c[i].m_nonlinear.V1),
                                pmt::from_double(c[i].m_nonlinear.V2),
                                pmt::from_long(c[i].m_nonlinear.p1),
                                pmt::from_long(c[i].m_nonlinear.p2));
        }
        pmt::vector_set(v, i, t);
    }
    return v;
}

struct cwd_guard {
    char old[4096];
    bool ok;
    explicit cwd_guard(const char *dir)
    {
        ok = getcwd(old, sizeof(old)) != NULL && dir && chdir(dir) == 0;
    }
    ~cwd_guard()
    {
        if (ok && chdir(old) != 0) perror("chdir");
    }
};

extern "C" {

int ref_candidate_size(void) { return (int)sizeof(candidate_t); }

/* ---------------- FDR ---------------- */
void *ref_fdr_new(int fs, int fl, int spb, int maxdrift, int maxfreqs, int halfbandwidth, int cf,
                  int threshold)
{
    return new FDR_impl(fs, fl, spb, maxdrift, maxfreqs, halfbandwidth, cf, threshold);
}
void ref_fdr_free(void *h) { delete (FDR_impl *)h; }

void ref_fdr_dims(void *h, int *n, int *size, int *hpbm, int *m, float *df, float *min_snr)
{
    FDR_impl *f = (FDR_impl *)h;
    if (n) *n = f->n;
    if (size) *size = f->size;
    if (hpbm) *hpbm = f->hpbm;
    if (m) *m = f->m;
    if (df) *df = f->df;
    if (min_snr) *min_snr = f->min_snr;
}

void ref_fdr_window(void *h, float *w) { memcpy(w, ((FDR_impl *)h)->w, sizeof(float) * ((FDR_impl *)h)->size); }

/* Runs FDR_impl::transform on one window.  iq: fl interleaved complex64.
 * cand_out: maxfreqs records of sizeof(candidate_t); ps_out [n][size], psavg_out [size] may be NULL.
 * spectra (may be NULL): [n][size][2] spectra that replace the stub FFT's output.
 * Returns npk as published in the output PDU; the PDU stays in the block's outbox
 * until ref_fdr_take_pdu()/ref_pipeline consumes it (it is dropped here otherwise). */
int ref_fdr_transform(void *h, const float *iq, void *cand_out, float *ps_out, float *psavg_out,
                      const float *spectra)
{
    FDR_impl *f = (FDR_impl *)h;
    inject_t inj = { spectra, f->n, 0 };
    if (spectra) oracle_fft_set_hook(inject_hook, &inj);
    pmt::pmt_t pdu = pmt::cons(pmt::PMT_NIL, samples_to_pmt(iq, f->fl));
    f->oracle_deliver("in", pdu);
    oracle_fft_set_hook(NULL, NULL);
    pmt::pmt_t out = f->oracle_outbox().back();
    f->oracle_outbox().clear();
    int npk = (int)pmt::to_long(pmt::tuple_ref(pmt::cdr(out), 1));
    if (cand_out) memcpy(cand_out, f->candidates, sizeof(candidate_t) * npk);
    if (ps_out)
        for (int i = 0; i < f->n; i++) memcpy(ps_out + (size_t)i * f->size, f->ps[i], sizeof(float) * f->size);
    if (psavg_out) memcpy(psavg_out, f->psavg, sizeof(float) * f->size);
    return npk;
}

/* ---------------- sync_and_demodulate ---------------- */
void *ref_sd_new(int fs, int fl, int spb, int maxdrift, int maxfreqs, int cf, const char *logdir)
{
    cwd_guard g(logdir ? logdir : "/tmp"); /* the ctor opens ./messagelog.txt for append */
    return new sync_and_demodulate_impl(fs, fl, spb, maxdrift, maxfreqs, cf);
}
void ref_sd_free(void *h) { delete (sync_and_demodulate_impl *)h; }

/* One direct call of sync_and_demodulate_impl::sync_and_demodulate (not traced). */
void ref_sd_eval(void *h, const void *cand, const float *id, const float *qd, long np,
                 unsigned char *symbols, float *f1, int ifmin, int ifmax, float fstep, int *shift1,
                 int lagmin, int lagmax, int lagstep, float *drift1, int symfac, float *sync,
                 int mode)
{
    sync_and_demodulate_impl *s = (sync_and_demodulate_impl *)h;
    ref_trace_t *saved = g_trace;
    g_trace = NULL;
    s->sync_and_demodulate(*(const candidate_t *)cand, (float *)id, (float *)qd, np, symbols, f1,
                           ifmin, ifmax, fstep, shift1, lagmin, lagmax, lagstep, drift1, symfac,
                           sync, mode);
    g_trace = saved;
}

static int run_demodulate(sync_and_demodulate_impl *s, pmt::pmt_t pdu, ref_trace_t *trace,
                          signed char *blobs, int max_blobs)
{
    if (trace) {
        trace->n_calls = 0;
        trace->n_fanos = 0;
    }
    g_trace = trace;
    s->oracle_deliver("in", pdu);
    g_trace = NULL;
    int nb = 0;
    for (pmt::pmt_t &m : s->oracle_outbox()) {
        pmt::pmt_t payload = pmt::cdr(m);
        if (blobs && nb < max_blobs) memcpy(blobs + 7 * nb, pmt::blob_data(payload), 7);
        nb++;
    }
    s->oracle_outbox().clear();
    return nb;
}

/* Runs sync_and_demodulate_impl::demodulate on one window with the given candidates.
 * Returns the number of 7-byte blobs published (one per decoded candidate). */
int ref_sd_demodulate(void *h, const float *iq, const void *cands, int npk, ref_trace_t *trace,
                      signed char *blobs, int max_blobs)
{
    sync_and_demodulate_impl *s = (sync_and_demodulate_impl *)h;
    pmt::pmt_t tuple = pmt::make_tuple(samples_to_pmt(iq, s->fl), pmt::from_long(npk),
                                       candidates_to_pmt((const candidate_t *)cands, npk));
    return run_demodulate(s, pmt::cons(pmt::PMT_NIL, tuple), trace, blobs, max_blobs);
}

/* FDR -> sync_and_demodulate exactly as the flowgraph wires them: the PDU the
 * FDR block publishes is handed to the demodulator untouched. */
int ref_pipeline(void *hf, void *hs, const float *iq, void *cand_out, int *npk_out,
                 ref_trace_t *trace, signed char *blobs, int max_blobs)
{
    FDR_impl *f = (FDR_impl *)hf;
    sync_and_demodulate_impl *s = (sync_and_demodulate_impl *)hs;
    pmt::pmt_t pdu = pmt::cons(pmt::PMT_NIL, samples_to_pmt(iq, f->fl));
    f->oracle_deliver("in", pdu);
    pmt::pmt_t out = f->oracle_outbox().back();
    f->oracle_outbox().clear();
    int npk = (int)pmt::to_long(pmt::tuple_ref(pmt::cdr(out), 1));
    if (npk_out) *npk_out = npk;
    if (cand_out) memcpy(cand_out, f->candidates, sizeof(candidate_t) * npk);
    return run_demodulate(s, out, trace, blobs, max_blobs);
}

void ref_fano_stats(double *seconds, long *calls, int reset)
{
    if (seconds) *seconds = g_fano_seconds;
    if (calls) *calls = g_fano_calls;
    if (reset) {
        g_fano_seconds = 0;
        g_fano_calls = 0;
    }
}
void ref_set_skip_fano(int skip) { g_skip_fano = skip; }

/* ---------------- helpers that live in the reference ---------------- */
float ref_slm_frequency_drift(double V1, double V2, int p1, int p2, float cf, float t)
{
    SLM slm;
    mode_nonlinear m;
    m.V1 = V1;
    m.V2 = V2;
    m.p1 = p1;
    m.p2 = p2;
    return slm.slmFrequencyDrift(m, cf, t);
}

/* enumerates the generator (slm.cc:76-116); out: [125][4] doubles V1,V2,p1,p2; returns count */
int ref_slm_generate(double *out, int max)
{
    SLM slm;
    mode_nonlinear m;
    int n = 0;
    slm.slmGeneratorInit();
    while (slm.slmGenerator(&m)) {
        if (n < max) {
            out[4 * n + 0] = m.V1;
            out[4 * n + 1] = m.V2;
            out[4 * n + 2] = m.p1;
            out[4 * n + 3] = m.p2;
        }
        n++;
    }
    return n;
}

int ref_fano_encode(unsigned char *symbols, unsigned char *data, unsigned int nbytes)
{
    Fano f;
    return f.encode(symbols, data, nbytes);
}

void ref_fano_mettab(int *out /* [2][256] */)
{
    Fano f;
    memcpy(out, f.mettab, sizeof(int) * 512);
}

int ref_fano_decode(unsigned char *data11, unsigned char *symbols162, unsigned int *metric,
                    unsigned int *cycles, unsigned int *maxnp, int delta, unsigned int maxcycles)
{
    Fano f;
    ref_trace_t *saved = g_trace;
    g_trace = NULL;
    int r = f.fano(metric, cycles, maxnp, data11, symbols162, 81, f.mettab, delta, maxcycles);
    g_trace = saved;
    return r;
}

void ref_deinterleave(void *hs, unsigned char *sym162)
{
    ((sync_and_demodulate_impl *)hs)->deinterleave(sym162);
}

void ref_pr3(unsigned char *out162)
{
    /* pr3[] is a file-static array inside both block TUs; read it through the
     * demodulator's view by demodulating nothing: simply copy from the header. */
#include "pr3.h"
    memcpy(out162, pr3, 162);
}

/* ---------------- unpacker (lib/helpers.cc:494-590) ---------------- */
/* hashtab: 32768*13 bytes of caller-owned state; call_loc_pow >= 32 bytes; callsign >= 16 bytes */
int ref_unpk(const signed char *message7, char *hashtab, char *call_loc_pow, char *callsign)
{
    helpers h;
    char msg[11];
    memset(msg, 0, sizeof(msg));
    memcpy(msg, message7, 7);
    return h.unpk_(msg, hashtab, call_loc_pow, callsign);
}
unsigned int ref_nhash(const void *key, size_t length, unsigned int initval)
{
    helpers h;
    return h.nhash(key, length, initval);
}

/* ---------------- sliding window ---------------- */
void *ref_sw_new(int fs, int fl, int shift, int C)
{
    return new sliding_window_stream_to_pdu_impl(fs, fl, shift, C);
}
void ref_sw_free(void *h) { delete (sliding_window_stream_to_pdu_impl *)h; }
/* feeds n complex64 items to work(); returns the number of PDUs emitted by this
 * call (0 or 1) and copies the window (fl interleaved complex64) to out if so. */
int ref_sw_work(void *h, const float *iq, int n, float *out)
{
    sliding_window_stream_to_pdu_impl *s = (sliding_window_stream_to_pdu_impl *)h;
    gr_vector_const_void_star in(1, (const void *)iq);
    gr_vector_void_star outs;
    s->work(n, in, outs);
    int npdu = 0;
    for (pmt::pmt_t &m : s->oracle_outbox()) {
        pmt::pmt_t v = pmt::cdr(m);
        if (out)
            for (int i = 0; i < s->fl; i++) {
                std::complex<double> c = pmt::to_complex(pmt::vector_ref(v, i));
                out[2 * i] = (float)c.real();
                out[2 * i + 1] = (float)c.imag();
            }
        npdu++;
    }
    s->oracle_outbox().clear();
    return npdu;
}

} // extern "C"
