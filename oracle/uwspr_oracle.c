/* uwspr_oracle.c -- CPU restatement of the gr-uwspr receive hot path.
 * TEST INFRASTRUCTURE (see uwspr_oracle.h for the rules and the parity status).
 *
 * Build: gcc -O2 -ffp-contract=off (oracle/Makefile).  Every floating-point
 * expression below keeps the operand types and the evaluation order of the
 * reference line it cites, because candidate selection is decided by strict
 * comparisons of fp32 sums.  Where the reference computes in double and stores to
 * float the cast is written out.
 */
#include "uwspr_oracle.h"
#include "uwspr_oracle_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int orc_sync_bit(int i) { return (int)((ORC_SYNC_WORDS[i >> 5] >> (i & 31)) & 1u); }

/* hazard H1 (sync_and_demodulate_impl.cc:177): 0 -> t == 0 as the compiled
 * reference behaves; 1 -> the apparently intended t = i*111/162 */
static int g_nonlinear_intended_t = 0;
void orc_set_nonlinear_intended_t(int on) { g_nonlinear_intended_t = on; }

/* ------------------------------------------------------------------ FDR ctor */
int orc_fdr_init(orc_fdr_t *f, int fs, int fl, int spb, int maxdrift, int maxfreqs,
                 int halfbandwidth, int cf, int threshold)
{
    memset(f, 0, sizeof(*f));
    f->fs = fs;
    f->fl = fl;
    f->spb = spb;
    f->maxdrift = maxdrift;
    f->maxfreqs = maxfreqs;
    f->halfbandwidth = halfbandwidth;
    f->cf = cf;
    f->threshold = (float)threshold;                      /* FDR_impl.cc:79 */
    f->size = 2 * spb;                                    /* :81 */
    if (f->size > 4096 || f->size < 8) return -1;
    int maxfreq = (int)((float)fs / 2.0);                 /* :82 */
    if (halfbandwidth > maxfreq) return -1;               /* :85-90 (reference exits) */
    f->df = (float)fs / (float)f->size;                   /* :93 */
    f->m = f->size / 2;                                   /* :95 */
    /* :97 -- both casts bind to `halfbandwidth` alone, the quotient is int/float */
    f->hpbm = (int)ceilf((float)(int)(float)halfbandwidth / f->df);
    for (int i = 0; i < f->size; i++)                     /* :103-105 */
        f->w[i] = (float)sin((M_PI / (f->size - 1)) * i);
    f->n = (int)(floor(((float)fl / (float)spb) * 2.0) - 3); /* :109 */
    f->min_snr = (float)pow(10.0, -7.0 / 10.0);           /* :137 */
    if (f->hpbm + 3 > f->m || f->n < 1) return -1;        /* hazard H3: psavg[m-hpbm-3 ..] */
    return 0;
}

/* -------------------------------------------------------- spectrogram */
/* forward DFT in double precision, radix-2 decimation in time (power-of-two n);
 * same definition as oracle/stubs/fftw3_stub.c, rounded once to float */
static void dft_forward(int n, const float *in /* [n][2] */, float *out /* [n][2] */)
{
    static __thread int cn = 0;
    static __thread double *wr = NULL, *wi = NULL, *re = NULL, *im = NULL;
    static __thread int *rev = NULL;
    if (cn != n) {
        free(wr);
        free(rev);
        wr = (double *)malloc(sizeof(double) * 3 * n);
        wi = wr + n / 2;
        re = wr + n;
        im = re + n;
        rev = (int *)malloc(sizeof(int) * n);
        int bits = 0;
        while ((1 << bits) < n) bits++;
        for (int i = 0; i < n; i++) {
            int r = 0;
            for (int b = 0; b < bits; b++)
                if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            rev[i] = r;
        }
        for (int j = 0; j < n / 2; j++) {
            double a = -2.0 * M_PI * j / n;
            wr[j] = cos(a);
            wi[j] = sin(a);
        }
        cn = n;
    }
    for (int i = 0; i < n; i++) {
        re[rev[i]] = in[2 * i];
        im[rev[i]] = in[2 * i + 1];
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int j = 0; j < half; j++) {
            double cr = wr[j * step], ci = wi[j * step];
            for (int s = 0; s < n; s += len) {
                int u = s + j, v = u + half;
                double tr = re[v] * cr - im[v] * ci;
                double ti = re[v] * ci + im[v] * cr;
                re[v] = re[u] - tr;
                im[v] = im[u] - ti;
                re[u] += tr;
                im[u] += ti;
            }
        }
    }
    for (int k = 0; k < n; k++) {
        out[2 * k] = (float)re[k];
        out[2 * k + 1] = (float)im[k];
    }
}

void orc_power(const orc_fdr_t *f, const float *spectra, float *ps)
{
    const int size = f->size;
    for (int i = 0; i < f->n; i++) {
        const float *X = spectra + (size_t)i * size * 2;
        for (int j = 0; j < size; j++) {                  /* FDR_impl.cc:246-253 */
            int k = j + f->spb;
            if (k > size - 1) k -= size;
            ps[(size_t)i * size + j] = X[2 * k] * X[2 * k] + X[2 * k + 1] * X[2 * k + 1];
        }
    }
}

void orc_spectrogram(const orc_fdr_t *f, const float *iq, float *ps, float *spectra_out)
{
    const int size = f->size;
    float *in = (float *)malloc(sizeof(float) * 4 * size);
    float *out = in + 2 * size;
    for (int i = 0; i < f->n; i++) {
        for (int j = 0; j < size; j++) {                  /* :224-232 */
            int k = i * (f->spb / 2) + j;
            /* the reference multiplies a double holding an fp32 sample by the fp32
             * window in double and stores to float: the correctly rounded fp32 product */
            in[2 * j] = (float)((double)iq[2 * k] * f->w[j]);
            in[2 * j + 1] = (float)((double)iq[2 * k + 1] * f->w[j]);
        }
        dft_forward(size, in, out);                       /* :244 */
        if (spectra_out) memcpy(spectra_out + (size_t)i * size * 2, out, sizeof(float) * 2 * size);
        for (int j = 0; j < size; j++) {                  /* :246-253 */
            int k = j + f->spb;
            if (k > size - 1) k -= size;
            ps[(size_t)i * size + j] = out[2 * k] * out[2 * k] + out[2 * k + 1] * out[2 * k + 1];
        }
    }
    free(in);
}

/* ------------------------------------------------ normalizer + peak pick */
static int float_less(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x < y) ? -1 : (x > y);
}

int orc_normalize_peaks(const orc_fdr_t *f, const float *ps, float *psavg_out, float *smspec_out,
                        orc_candidate_t *cands)
{
    const int size = f->size, hpbm = f->hpbm, m = f->m;
    const int finpb = 2 * hpbm;                           /* FDR_impl.cc:265 */
    float *psavg = (float *)malloc(sizeof(float) * (size + 2 * finpb));
    float *smspec = psavg + size, *tmpsort = smspec + finpb;

    for (int j = 0; j < size; j++) {                      /* :257-263, row order */
        float acc = 0;
        for (int i = 0; i < f->n; i++) acc = acc + ps[(size_t)i * size + j];
        psavg[j] = acc;
    }
    for (int i = 0; i < finpb; i++) {                     /* :268-275 */
        float acc = 0.0f;
        for (int j = -3; j <= 3; j++) acc = acc + psavg[m - hpbm + i + j];
        smspec[i] = acc;
    }
    memcpy(tmpsort, smspec, sizeof(float) * finpb);       /* :277-281 */
    qsort(tmpsort, finpb, sizeof(float), float_less);
    int noiseidx = (int)floor(0.3 * (float)finpb);        /* :283 */
    float noise_level = tmpsort[noiseidx];
    const float floor_val = (float)(0.1 * f->min_snr);    /* :290, double product stored to float */
    for (int j = 0; j < finpb; j++) {                     /* :287-291 */
        float q = smspec[j] / noise_level;
        smspec[j] = (float)((double)q - 1.0);
        if (smspec[j] < f->min_snr) smspec[j] = floor_val;
    }
    int npk = 0;
    for (int j = 1; j < finpb - 1; j++) {                 /* :294-306 */
        if (smspec[j] > smspec[j - 1] && smspec[j] > smspec[j + 1] && npk < f->maxfreqs) {
            memset(&cands[npk], 0, sizeof(cands[npk]));
            cands[npk].freq = (j - hpbm) * f->df;
            cands[npk].snr = 10 * log10f(smspec[j]);      /* float overload of log10 (C++ <math.h>) */
            npk++;
        }
    }
    /* :311-319 -- stable descending exchange sort on snr */
    for (int pass = 1; pass <= npk - 1; pass++)
        for (int k = 0; k < npk - pass; k++)
            if (cands[k].snr < cands[k + 1].snr) {
                orc_candidate_t t = cands[k];
                cands[k] = cands[k + 1];
                cands[k + 1] = t;
            }
    if (psavg_out) memcpy(psavg_out, psavg, sizeof(float) * size);
    if (smspec_out) memcpy(smspec_out, smspec, sizeof(float) * finpb);
    free(psavg);
    return npk;
}

/* ------------------------------------------------------------------- SLM */
float orc_slm_frequency_drift(double V1, double V2, int p1, int p2, float cf, float t)
{
    const float c = 1500.0f;                              /* slm.cc:40 */
    double q1 = V1 * t + p1, q2 = V2 * t + p2;            /* connecting vector */
    float sign = (float)(((q1 * V1 + q2 * V2) > 0) * 2 - 1); /* :42-54 */
    double numerator = fabs(V1 * q1 + V2 * q2);           /* :56-61 */
    double denominator = sqrt(q1 * q1 + q2 * q2);         /* :63-67, pow(x,2) == x*x */
    if (denominator == 0) return 0.0f;
    return (float)(-sign * numerator / denominator * cf / c); /* :72 */
}

int orc_slm_trajectory(int k, double *V1, double *V2, int *p1, int *p2)
{
    /* slm.cc:76-116: p2 index fastest, then V1, then V2; 5 x 5 x 5 instances */
    if (k < 0 || k >= 125) return 0;
    int ip2 = k % 5, iV1 = (k / 5) % 5, iV2 = k / 25;
    *V1 = iV1 * 1.0 + -2.0;
    *V2 = iV2 * 1.0 + -2.0;
    *p1 = 0;
    *p2 = ip2 * 200 + 50;
    return 1;
}

/* ---------------------------------------------------------- coarse search */
/* lib/FDR_impl.cc:188-210 */
static inline void powersum(const float *ps, int size, int k0, int k, int ifd, float *ss, float *pw)
{
    const float *row = ps + (size_t)(k0 + 2 * k) * size;
    float p0 = sqrtf(row[ifd - 3]);
    float p1 = sqrtf(row[ifd - 1]);
    float p2 = sqrtf(row[ifd + 1]);
    float p3 = sqrtf(row[ifd + 3]);
    *ss = *ss + (float)(2 * orc_sync_bit(k) - 1) * ((p1 + p3) - (p0 + p2));
    *pw = *pw + p0 + p1 + p2 + p3;
}

void orc_coarse(const orc_fdr_t *f, const float *ps, orc_candidate_t *cands, int npk)
{
    const int size = f->size, m = f->m;
    const float df = f->df;
    /* drift of every trajectory at every integer second the search visits (:382-385) */
    static __thread float slm_tab[125][162];
    static __thread int slm_cf = -1;
    if (slm_cf != f->cf) {
        for (int h = 0; h < 125; h++) {
            double V1, V2;
            int p1, p2;
            orc_slm_trajectory(h, &V1, &V2, &p1, &p2);
            for (int k = 0; k < 162; k++) {
                float t = (float)(k * 111 / 162);         /* integer division, :382 */
                slm_tab[h][k] = orc_slm_frequency_drift(V1, V2, p1, p2, (float)f->cf, t);
            }
        }
        slm_cf = f->cf;
    }
    for (int j = 0; j < npk; j++) {
        orc_candidate_t *c = &cands[j];
        c->sync = -1e30f;                                 /* :340 */
        int if0 = (int)(c->freq / df + m);                /* :341 */
        for (int ifr = if0 - 2; ifr <= if0 + 2; ifr++) {
            for (int k0 = 0; k0 < 26; k0++) {
                for (int drift = -f->maxdrift; drift <= f->maxdrift; drift++) { /* :348-374 */
                    float ss = 0.0f, pw = 0.0f;
                    for (int k = 0; k < 162; k++) {
                        int ifd = (int)(ifr + ((float)k - 81.0) / 81.0 * ((float)drift) / (2.0 * df)); /* :353 */
                        powersum(ps, size, k0, k, ifd, &ss, &pw);
                    }
                    float sync = ss / pw;
                    if (sync > c->sync) {                 /* :360 */
                        c->shift = 128 * k0;
                        c->freq = (ifr - m) * df;
                        c->sync = sync;
                        c->m_type = 0;
                        c->u.lin.drift = (float)drift;
                    }
                }
                for (int h = 0; h < 125; h++) {           /* :376-405 */
                    float ss = 0.0f, pw = 0.0f;
                    for (int k = 0; k < 162; k++) {
                        int ifd = (int)(ifr + slm_tab[h][k] / df); /* :384-385, fp32 */
                        powersum(ps, size, k0, k, ifd, &ss, &pw);
                    }
                    float sync = ss / pw;
                    if (sync / c->sync > f->threshold) {  /* :392 */
                        c->shift = 128 * k0;
                        c->freq = (ifr - m) * df;
                        c->sync = sync;
                        c->m_type = 1;
                        orc_slm_trajectory(h, &c->u.nl.V1, &c->u.nl.V2, &c->u.nl.p1, &c->u.nl.p2);
                    }
                }
            }
        }
    }
}

int orc_fdr_transform(const orc_fdr_t *f, const float *iq, orc_candidate_t *cands, float *ps_scratch)
{
    float *ps = ps_scratch ? ps_scratch : (float *)malloc(sizeof(float) * (size_t)f->n * f->size);
    orc_spectrogram(f, iq, ps, NULL);
    int npk = orc_normalize_peaks(f, ps, NULL, NULL, cands);
    orc_coarse(f, ps, cands, npk);
    if (!ps_scratch) free(ps);
    return npk;
}

/* ------------------------------------------------- fine sync / soft symbols */
void orc_sync_and_demodulate(const orc_candidate_t *cand, int cf, const float *id, const float *qd,
                             long np, unsigned char *symbols, float *f1, int ifmin, int ifmax,
                             float fstep, int *shift1, int lagmin, int lagmax, int lagstep,
                             float *drift1, int symfac, float *sync, int mode)
{
    const float dt = (float)(1.0 / 375.0), df = (float)(375.0 / 256.0); /* :146 */
    const float delta[4] = { (float)(-df * 1.5), (float)(-df * 0.5), (float)(df * 0.5), (float)(df * 1.5) }; /* :148 */
    float c[4][256], s[4][256];
    float fsymb[162];
    float syncmax = -1e30f, f0 = 0.0f, fbest = 0.0f, fplast = -10000.0f;
    int best_shift = 0;
    memset(fsymb, 0, sizeof(fsymb));

    if (mode == 0) { ifmin = 0; ifmax = 0; fstep = 0.0f; }                  /* :160 */
    if (mode == 1) { lagmin = *shift1; lagmax = *shift1; }                  /* :161 */
    if (mode == 2) { lagmin = *shift1; lagmax = *shift1; ifmin = 0; ifmax = 0; } /* :162 */
    f0 = *f1;
    if (lagstep <= 0) lagstep = 1; /* hazard H4: the reference would never return */

    for (int ifreq = ifmin; ifreq <= ifmax; ifreq++) {
        f0 = *f1 + ifreq * fstep;                                           /* :164 */
        for (int lag = lagmin; lag <= lagmax; lag += lagstep) {
            float ss = 0.0f, totp = 0.0f;
            for (int i = 0; i < 162; i++) {
                float fp;
                if (cand->m_type == 0) {
                    fp = (float)(f0 + (*drift1 / 2.0) * ((float)i - 81.0) / 81.0); /* :173 */
                } else {
                    /* :177-180: `t` is never assigned (the assignment sits between a
                     * break and the next case label); the compiled reference reads 0 */
                    float t = g_nonlinear_intended_t ? (float)(i * 111 / 162) : 0.0f;
                    fp = f0 + orc_slm_frequency_drift(cand->u.nl.V1, cand->u.nl.V2, cand->u.nl.p1,
                                                      cand->u.nl.p2, (float)cf, t);
                }
                if (i == 0 || fp != fplast) {                               /* :185-199 */
                    for (int j = 0; j < 4; j++) {
                        float cdphi = (float)cos(2 * M_PI * dt * (fp + delta[j]));
                        float sdphi = (float)sin(2 * M_PI * dt * (fp + delta[j]));
                        c[j][0] = 1;
                        s[j][0] = 0;
                        for (int k = 1; k < 256; k++) {
                            c[j][k] = c[j][k - 1] * cdphi - s[j][k - 1] * sdphi;
                            s[j][k] = c[j][k - 1] * sdphi + s[j][k - 1] * cdphi;
                        }
                    }
                    fplast = fp;
                }
                float p[4];
                for (int j = 0; j < 4; j++) {                               /* :200-212 */
                    float inp = 0.0f, quad = 0.0f;
                    for (int k = 0; k < 256; k++) {
                        long n = (long)lag + i * 256 + k;
                        if (n > 0 && n < np) {                              /* hazard H5: n == 0 skipped */
                            inp = inp + id[n] * c[j][k] + qd[n] * s[j][k];
                            quad = quad - id[n] * s[j][k] + qd[n] * c[j][k];
                        }
                    }
                    p[j] = sqrtf(inp * inp + quad * quad);
                }
                totp = totp + p[0] + p[1] + p[2] + p[3];                    /* :213 */
                float cmet = (p[1] + p[3]) - (p[0] + p[2]);                 /* :214 */
                ss = orc_sync_bit(i) ? ss + cmet : ss - cmet;               /* :215 */
                if (mode == 2)                                              /* :216-224 */
                    fsymb[i] = orc_sync_bit(i) ? p[3] - p[1] : p[2] - p[0];
            }
            ss = ss / totp;                                                 /* :226 */
            if (ss > syncmax) {
                syncmax = ss;
                best_shift = lag;
                fbest = f0;
            }
        }
    }
    if (mode <= 1) {                                                        /* :234-239 */
        *sync = syncmax;
        *shift1 = best_shift;
        *f1 = fbest;
        return;
    }
    if (mode == 2) {                                                        /* :240-254 */
        float fsum = 0.0f, f2sum = 0.0f;
        *sync = syncmax;
        for (int i = 0; i < 162; i++) {
            fsum = (float)(fsum + fsymb[i] / 162.0);
            f2sum = (float)(f2sum + fsymb[i] * fsymb[i] / 162.0);
        }
        float fac = sqrtf(f2sum - fsum * fsum);
        for (int i = 0; i < 162; i++) {
            float v = symfac * fsymb[i] / fac;
            if (v > 127) v = 127.0f;
            if (v < -128) v = -128.0f;
            float q = v + 128;
            /* float -> unsigned char of a NaN is undefined in C; x86 yields 0 */
            symbols[i] = (q == q) ? (unsigned char)q : 0;
        }
    }
}

/* inverse of the bit-reversal order used by :265-282 */
static void bitrev_order(unsigned char *order /* [162] */)
{
    int p = 0;
    for (int i = 0; p < 162 && i < 256; i++) {
        int j = 0;
        for (int b = 0; b < 8; b++)
            if (i & (1 << b)) j |= 1 << (7 - b);
        if (j < 162) order[p++] = (unsigned char)j;
    }
}

void orc_deinterleave(unsigned char *sym)
{
    unsigned char order[162], tmp[162];
    bitrev_order(order);
    for (int p = 0; p < 162; p++) tmp[p] = sym[order[p]];
    memcpy(sym, tmp, 162);
}

void orc_interleave(unsigned char *sym)
{
    unsigned char order[162], tmp[162];
    bitrev_order(order);
    for (int p = 0; p < 162; p++) tmp[order[p]] = sym[p];
    memcpy(sym, tmp, 162);
}

static void trace_call(orc_trace_t *tr, const orc_sd_call_t *rec)
{
    if (!tr) return;
    if (tr->calls && tr->n_calls < tr->max_calls) tr->calls[tr->n_calls] = *rec;
    tr->n_calls++;
}

/* one traced refinement call */
static void sd_call(orc_trace_t *tr, const orc_candidate_t *cand, int cf, const float *id,
                    const float *qd, unsigned char *symbols, float *f1, int ifmin, int ifmax,
                    float fstep, int *shift1, int lagmin, int lagmax, int lagstep, float *drift1,
                    int symfac, float *sync, int mode)
{
    orc_sd_call_t rec;
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
    orc_sync_and_demodulate(cand, cf, id, qd, 45000 /* :92 */, symbols, f1, ifmin, ifmax, fstep,
                            shift1, lagmin, lagmax, lagstep, drift1, symfac, sync, mode);
    rec.f1_out = *f1;
    rec.shift_out = *shift1;
    rec.sync_out = *sync;
    if (mode == 2) memcpy(rec.symbols, symbols, 162);
    trace_call(tr, &rec);
}

int orc_demodulate_ex(int cf, const float *iq, int fl, const orc_candidate_t *cands_in, int npk,
                      orc_trace_t *trace, unsigned char *blobs, int max_blobs, int run_fano,
                      float *refined /* [npk][5]: f1, shift1, drift1, sync1, worth_a_try at :457 */);

int orc_demodulate(int cf, const float *iq, int fl, const orc_candidate_t *cands_in, int npk,
                   orc_trace_t *trace, unsigned char *blobs, int max_blobs, int run_fano)
{
    return orc_demodulate_ex(cf, iq, fl, cands_in, npk, trace, blobs, max_blobs, run_fano, NULL);
}

int orc_demodulate_ex(int cf, const float *iq, int fl, const orc_candidate_t *cands_in, int npk,
                      orc_trace_t *trace, unsigned char *blobs, int max_blobs, int run_fano,
                      float *refined)
{
    /* tuning constants, sync_and_demodulate_impl.cc:326-335 */
    const unsigned int maxcycles = 10000;
    const float minsync1 = 0.10f, minsync2 = 0.12f;
    const int iifac = 8, symfac = 50, delta = 60;
    const float minrms = (float)(52.0 * (symfac / 64.0));
    int nblobs = 0;
    float *idat = (float *)malloc(sizeof(float) * 2 * (size_t)fl);
    float *qdat = idat + fl;
    unsigned char symbols[162], decdata[11];
    for (int i = 0; i < fl; i++) {                                          /* :341-346 */
        idat[i] = iq[2 * i];
        qdat[i] = iq[2 * i + 1];
    }
    if (trace) {
        trace->n_calls = 0;
        trace->n_fanos = 0;
    }
    for (int j = 0; j < npk; j++) {                                         /* :389 */
        orc_candidate_t cand = cands_in[j];
        float drift_in = (cand.m_type == 0) ? cand.u.lin.drift : 0.0f;      /* :360,:373 */
        if (cand.m_type != 0) {
            /* :373 stores 0.0f over the low half of V1 in the union */
            float z = 0.0f;
            memcpy(&cand.u, &z, sizeof(z));
        }
        memset(symbols, 0, sizeof(symbols));
        float f1 = cand.freq, drift1 = drift_in, sync1 = cand.sync;
        int shift1 = cand.shift;
        float fstep = 0.0f;
        int ifmin = 0, ifmax = 0;
        int lagmin = shift1 - 128, lagmax = shift1 + 128, lagstep = 64;     /* :409-412 */
        sd_call(trace, &cand, cf, idat, qdat, symbols, &f1, ifmin, ifmax, fstep, &shift1, lagmin,
                lagmax, lagstep, &drift1, symfac, &sync1, 0);
        fstep = 0.25f; ifmin = -2; ifmax = 2;                               /* :416 */
        sd_call(trace, &cand, cf, idat, qdat, symbols, &f1, ifmin, ifmax, fstep, &shift1, lagmin,
                lagmax, lagstep, &drift1, symfac, &sync1, 1);
        if (cand.m_type == 0) {                                             /* :423-441 */
            fstep = 0.0f; ifmin = 0; ifmax = 0;
            float driftp = (float)(drift1 + 0.5), driftm, syncp, syncm;
            sd_call(trace, &cand, cf, idat, qdat, symbols, &f1, ifmin, ifmax, fstep, &shift1,
                    lagmin, lagmax, lagstep, &driftp, symfac, &syncp, 1);
            driftm = (float)(drift1 - 0.5);
            sd_call(trace, &cand, cf, idat, qdat, symbols, &f1, ifmin, ifmax, fstep, &shift1,
                    lagmin, lagmax, lagstep, &driftm, symfac, &syncm, 1);
            if (syncp > sync1) {
                drift1 = driftp;
                sync1 = syncp;
            } else if (syncm > sync1) {
                drift1 = driftm;
                sync1 = syncm;
            }
        }
        int worth_a_try;
        if (sync1 > minsync1) {                                             /* :443-456 */
            lagmin = shift1 - 32; lagmax = shift1 + 32; lagstep = 16;
            sd_call(trace, &cand, cf, idat, qdat, symbols, &f1, ifmin, ifmax, fstep, &shift1,
                    lagmin, lagmax, lagstep, &drift1, symfac, &sync1, 0);
            fstep = 0.05f; ifmin = -2; ifmax = 2;
            sd_call(trace, &cand, cf, idat, qdat, symbols, &f1, ifmin, ifmax, fstep, &shift1,
                    lagmin, lagmax, lagstep, &drift1, symfac, &sync1, 1);
            worth_a_try = 1;
        } else {
            worth_a_try = 0;
        }
        if (refined) {
            refined[5 * j + 0] = f1;
            refined[5 * j + 1] = (float)shift1;
            refined[5 * j + 2] = drift1;
            refined[5 * j + 3] = sync1;
            refined[5 * j + 4] = (float)worth_a_try;
        }
        int idt = 0, not_decoded = 1;
        while (worth_a_try && not_decoded && idt <= (128 / iifac)) {        /* :460-482 */
            int ii = (idt + 1) / 2;
            if (idt % 2 == 1) ii = -ii;
            ii = iifac * ii;
            int jiggered_shift = shift1 + ii;
            sd_call(trace, &cand, cf, idat, qdat, symbols, &f1, ifmin, ifmax, fstep,
                    &jiggered_shift, lagmin, lagmax, lagstep, &drift1, symfac, &sync1, 2);
            float sq = 0.0f;
            for (int i = 0; i < 162; i++) {
                float y = (float)((float)symbols[i] - 128.0);
                sq += y * y;
            }
            float rms = (float)sqrt(sq / 162.0);
            if (sync1 > minsync2 && rms > minrms) {
                orc_deinterleave(symbols);
                orc_fano_call_t fr;
                memset(&fr, 0, sizeof(fr));
                memcpy(fr.symbols, symbols, 162);
                if (run_fano) {
                    not_decoded = orc_fano(&fr.metric, &fr.cycles, &fr.maxnp, decdata, symbols, 81,
                                           delta, maxcycles);
                    memcpy(fr.data, decdata, 11);
                } else {
                    not_decoded = -1;
                }
                fr.result = not_decoded;
                if (trace) {
                    if (trace->fanos && trace->n_fanos < trace->max_fanos) trace->fanos[trace->n_fanos] = fr;
                    trace->n_fanos++;
                }
            }
            idt++;
        }
        if (worth_a_try && !not_decoded) {                                  /* :483-531 */
            if (blobs && nblobs < max_blobs) memcpy(blobs + 7 * nblobs, decdata, 7);
            nblobs++;
        }
    }
    free(idat);
    return nblobs;
}

/* ------------------------------------------------------- convolutional code */
#define ORC_POLY1 0xf2d05351u /* Layland-Lushbaugh K=32 r=1/2, Fano.cc:54-55 */
#define ORC_POLY2 0xe4613c47u

static inline unsigned parity32(unsigned v)
{
    v ^= v >> 16;
    v ^= v >> 8;
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    return v & 1u;
}
/* symbol pair for an encoder state: POLY1 parity in bit 1, POLY2 parity in bit 0 (Fano.cc:63-72) */
static inline unsigned branch_symbol(unsigned long state)
{
    unsigned st = (unsigned)state;
    return (parity32(st & ORC_POLY1) << 1) | parity32(st & ORC_POLY2);
}

void orc_encode(unsigned char *symbols, const unsigned char *data, unsigned int nbytes)
{
    unsigned long state = 0;                              /* Fano.cc:81-100 */
    for (unsigned b = 0; b < nbytes; b++)
        for (int i = 7; i >= 0; i--) {
            state = (state << 1) | ((data[b] >> i) & 1u);
            unsigned sym = branch_symbol(state);
            *symbols++ = (unsigned char)(sym >> 1);
            *symbols++ = (unsigned char)(sym & 1u);
        }
}

void orc_channel_symbols(const unsigned char *msg7, unsigned char *chan162)
{
    unsigned char data[11], enc[176];
    memset(data, 0, sizeof(data));
    memcpy(data, msg7, 7);
    orc_encode(enc, data, 11);
    orc_interleave(enc); /* first 162 of the 176 encoder outputs */
    for (int i = 0; i < 162; i++) chan162[i] = (unsigned char)(2 * enc[i] + orc_sync_bit(i));
}

/* Fano sequential decoder, lib/Fano.cc:110-252 (algorithm of P. Karn, KA9Q).
 * Restated with array indices; cycle accounting identical to the reference so
 * that metric / cycles / maxnp can be compared. */
typedef struct {
    unsigned long encstate;
    long gamma;
    int metrics[4];
    int tm[2];
    int i;
} orc_node_t;

int orc_fano(unsigned int *metric, unsigned int *cycles, unsigned int *maxnp, unsigned char *data,
             const unsigned char *symbols, unsigned int nbits, int delta, unsigned int maxcycles)
{
    orc_node_t *nodes = (orc_node_t *)calloc(nbits + 1, sizeof(orc_node_t));
    const int last = (int)nbits - 1, tail = (int)nbits - 31;
    int np = 0;
    *maxnp = 0;
    for (int k = 0; k <= last; k++) {                     /* :140-147 */
        int a = symbols[2 * k], b = symbols[2 * k + 1];
        nodes[k].metrics[0] = ORC_METTAB[0][a] + ORC_METTAB[0][b];
        nodes[k].metrics[1] = ORC_METTAB[0][a] + ORC_METTAB[1][b];
        nodes[k].metrics[2] = ORC_METTAB[1][a] + ORC_METTAB[0][b];
        nodes[k].metrics[3] = ORC_METTAB[1][a] + ORC_METTAB[1][b];
    }
    nodes[0].encstate = 0;
    unsigned lsym = branch_symbol(nodes[0].encstate);     /* :150-168 */
    int m0 = nodes[0].metrics[lsym], m1 = nodes[0].metrics[3 ^ lsym];
    if (m0 > m1) {
        nodes[0].tm[0] = m0;
        nodes[0].tm[1] = m1;
    } else {
        nodes[0].tm[0] = m1;
        nodes[0].tm[1] = m0;
        nodes[0].encstate++;
    }
    nodes[0].i = 0;
    maxcycles *= nbits;
    nodes[0].gamma = 0;
    int t = 0;
    unsigned int i;
    for (i = 1; i <= maxcycles; i++) {                    /* :173-239 */
        if (np > (int)*maxnp) *maxnp = (unsigned)np;
        long ngamma = nodes[np].gamma + nodes[np].tm[nodes[np].i];
        if (ngamma >= t) {
            if (nodes[np].gamma < t + delta)
                while (ngamma >= t + delta) t += delta;
            nodes[np + 1].gamma = ngamma;
            nodes[np + 1].encstate = nodes[np].encstate << 1;
            if (++np == last + 1) break;
            lsym = branch_symbol(nodes[np].encstate);
            if (np >= tail) {
                nodes[np].tm[0] = nodes[np].metrics[lsym];
            } else {
                m0 = nodes[np].metrics[lsym];
                m1 = nodes[np].metrics[3 ^ lsym];
                if (m0 > m1) {
                    nodes[np].tm[0] = m0;
                    nodes[np].tm[1] = m1;
                } else {
                    nodes[np].tm[0] = m1;
                    nodes[np].tm[1] = m0;
                    nodes[np].encstate++;
                }
            }
            nodes[np].i = 0;
            continue;
        }
        for (;;) {                                        /* look backward, :217-238 */
            if (np == 0 || nodes[np - 1].gamma < t) {
                t -= delta;
                if (nodes[np].i != 0) {
                    nodes[np].i = 0;
                    nodes[np].encstate ^= 1;
                }
                break;
            }
            if (--np < tail && nodes[np].i != 1) {
                nodes[np].i++;
                nodes[np].encstate ^= 1;
                break;
            }
        }
    }
    *metric = (unsigned int)nodes[np].gamma;              /* :240 */
    unsigned nbytes = nbits >> 3;
    for (unsigned b = 0; b < nbytes; b++) data[b] = (unsigned char)nodes[7 + 8 * b].encstate;
    *cycles = i + 1;
    free(nodes);
    return (i >= maxcycles) ? -1 : 0;
}

long orc_sliding_window_count(long nitems, int chunk, int fs, int fl, int shift)
{
    /* sliding_window_stream_to_pdu_impl.cc:108-135 */
    long count = 0, windows = 0, left = nitems;
    while (left > 0) {
        long c = left < chunk ? left : chunk;
        count += c;
        left -= c;
        if (count >= fl) {
            windows++;
            count -= (long)shift * fs;
        }
    }
    return windows;
}
