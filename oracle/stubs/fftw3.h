/* Oracle stub for <fftw3.h> (test infrastructure, not product code).
 *
 * FFTW3f is a third-party dependency that is not under /root/reference and not
 * installed here, and the reference pins neither its version nor any FFT output
 * (lib/FDR_impl.cc:123-132,244 are the only call sites).  The stub therefore
 * implements the published definition of the forward DFT,
 *     X[k] = sum_n x[n] exp(-2 pi i n k / N),
 * evaluated in double precision (radix-2 decimation in time) and rounded once to
 * single precision -- the value any correct fp32 FFT approximates to ~1e-7.
 *
 * oracle_fft_set_hook() lets a harness substitute the spectrum of each execute
 * call (used by the parity tests to run the reference's own post-FFT code on
 * spectra produced by the CUDA path). */
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <stdio.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
struct oracle_fftwf_plan_s {
    int n;
    int sign;
    fftwf_complex *in, *out;
};
typedef struct oracle_fftwf_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_import_wisdom_from_file(FILE *f);
/* hook(user, n, in, out) returns non-zero if it filled `out` itself */
typedef int (*oracle_fft_hook_t)(void *user, int n, const fftwf_complex *in, fftwf_complex *out);
void oracle_fft_set_hook(oracle_fft_hook_t hook, void *user);
#ifdef __cplusplus
}
#endif
#endif
