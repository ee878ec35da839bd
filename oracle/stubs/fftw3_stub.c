/* Oracle stub FFT (test infrastructure): see fftw3.h in this directory. */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static oracle_fft_hook_t g_hook = NULL;
static void *g_hook_user = NULL;

void oracle_fft_set_hook(oracle_fft_hook_t hook, void *user)
{
    g_hook = hook;
    g_hook_user = user;
}

void *fftwf_malloc(size_t n)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64) != 0) return NULL;
    return p;
}
void fftwf_free(void *p) { free(p); }

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->n = n;
    p->sign = sign;
    p->in = in;
    p->out = out;
    return p;
}
void fftwf_destroy_plan(fftwf_plan p) { free(p); }
int fftwf_import_wisdom_from_file(FILE *f) { (void)f; return 0; }

/* double-precision DFT: radix-2 when n is a power of two, direct sum otherwise */
static void dft_double(int n, int sign, const fftwf_complex *in, double *re, double *im)
{
    int pow2 = n > 0 && (n & (n - 1)) == 0;
    if (!pow2) {
        for (int k = 0; k < n; k++) {
            double sr = 0, si = 0;
            for (int j = 0; j < n; j++) {
                double a = sign * 2.0 * M_PI * (double)(((long)j * k) % n) / n;
                double c = cos(a), s = sin(a);
                sr += in[j][0] * c - in[j][1] * s;
                si += in[j][0] * s + in[j][1] * c;
            }
            re[k] = sr;
            im[k] = si;
        }
        return;
    }
    /* twiddles exp(sign*2*pi*i*j/n) and the bit-reversal permutation are cached per (n, sign) */
    static __thread int c_n = 0, c_sign = 0;
    static __thread double *c_wr = NULL, *c_wi = NULL;
    static __thread int *c_rev = NULL;
    if (c_n != n || c_sign != sign) {
        free(c_wr);
        free(c_rev);
        c_wr = (double *)malloc(sizeof(double) * n);
        c_wi = c_wr + n / 2;
        c_rev = (int *)malloc(sizeof(int) * n);
        int bits = 0;
        while ((1 << bits) < n) bits++;
        for (int i = 0; i < n; i++) {
            int r = 0;
            for (int b = 0; b < bits; b++)
                if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            c_rev[i] = r;
        }
        for (int j = 0; j < n / 2; j++) {
            double a = sign * 2.0 * M_PI * j / n;
            c_wr[j] = cos(a);
            c_wi[j] = sin(a);
        }
        c_n = n;
        c_sign = sign;
    }
    for (int i = 0; i < n; i++) {
        re[c_rev[i]] = in[i][0];
        im[c_rev[i]] = in[i][1];
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int j = 0; j < half; j++) {
            double wr = c_wr[j * step], wi = c_wi[j * step];
            for (int s = 0; s < n; s += len) {
                int u = s + j, v = u + half;
                double tr = re[v] * wr - im[v] * wi;
                double ti = re[v] * wi + im[v] * wr;
                re[v] = re[u] - tr;
                im[v] = im[u] - ti;
                re[u] += tr;
                im[u] += ti;
            }
        }
    }
}

void fftwf_execute(const fftwf_plan p)
{
    if (g_hook && g_hook(g_hook_user, p->n, (const fftwf_complex *)p->in, p->out)) return;
    static __thread double *buf = NULL;
    static __thread int buf_n = 0;
    if (buf_n < p->n) {
        free(buf);
        buf = (double *)malloc(sizeof(double) * 2 * p->n);
        buf_n = p->n;
    }
    double *re = buf, *im = buf + p->n;
    dft_double(p->n, p->sign, (const fftwf_complex *)p->in, re, im);
    for (int k = 0; k < p->n; k++) {
        p->out[k][0] = (float)re[k];
        p->out[k][1] = (float)im[k];
    }
}
