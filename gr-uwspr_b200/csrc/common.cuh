// Shared declarations of the uwspr_b200 CUDA path (sm_100a).
//
// Arithmetic contract: the reference decides candidates by strict comparisons of fp32
// sums, and its CPU build cannot contract a*b+c into an FMA (x86-64 baseline).  All
// translation units are therefore compiled with -fmad=false and every expression that
// mirrors a reference line is written with the round-to-nearest intrinsics
// (__fadd_rn, __fmul_rn, __fdiv_rn, __fsqrt_rn) in the reference's evaluation order.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "uwspr_b200.h"

#define UW_NSYM 162
#define UW_FFT_N 512
#define UW_HOP 128
#define UW_NK0 26          // half-symbol start offsets searched, FDR_impl.cc:346
#define UW_NIFR 5          // if0-2 .. if0+2, FDR_impl.cc:344
#define UW_NTRAJ 125       // slm.cc:76-116
#define UW_SPS 256         // samples per symbol in the fine stage (literal in the reference)
#define UW_NP 45000        // npoints, sync_and_demodulate_impl.cc:92
#define UW_MAX_TILE_W 40   // widest amplitude tile (bins) one coarse CTA stages
#define UW_MAX_UNIQUE 192  // most distinct bin-offset sequences supported
#define UW_NQUAD 41        // ceil(162/4) packed offset words per sequence

// Scalar description of one context, passed by value to the kernels.
struct UwDims {
    int fl, n_rows, size, m, hpbm, finpb, noiseidx, maxfreqs, maxcand, maxdrift, cf;
    float df, min_snr, floor_val, threshold;
    int bin_lo, n_bins, nbp;        // kept bins [bin_lo, bin_lo+n_bins), row pitch nbp floats
    int n_lin, n_hyp, n_unique;     // 2*maxdrift+1, n_lin+125, distinct offset sequences
    int off_min, off_max, tile_w;   // bin offsets span [off_min, off_max]; tile_w = 11 + off_max - off_min
    int nonlinear_intended_t;
    uint32_t sync_words[6];         // WSPR sync vector, bit i at word i/32 bit i%32 (reference lib/pr3.h)
};

// one entry of the compact work list: which window, which candidate slot of that window
struct UwItem {
    int win, slot;
};

// per-window staging record written by the spectrogram kernel (before the coarse search)
struct UwPeak {
    float freq, snr;
};

__device__ __forceinline__ int uw_sync_bit(const uint32_t *words, int i)
{
    return (int)((words[i >> 5] >> (i & 31)) & 1u);
}

// ---- packed fp32 pairs (sm_100a: add/sub/fma.rn.f32x2 -> FADD2 / FFMA2) ------------------
// One issue slot carries two independent IEEE round-to-nearest operations, each bit-identical
// to its scalar form.  The kernels that mirror reference sums are bound by issue slots and
// the fp32 pipe (no FMA contraction allowed), so they pair two accumulation chains that
// share an operand (two tones against one sample, two bins against one shift).
// ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 even under
// --fmad=false, which would round once instead of twice.  The product is therefore issued
// as fma(a, b, -0.0) with the -0.0 pair in a kernel parameter (opaque to ptxas):
// a*b + (-0) is a*b rounded once, sign of zero included, and an FFMA2 cannot be merged
// with the addition that consumes it.
typedef unsigned long long uw_f2;
#define UW_NEGZERO2 0x8000000080000000ull

__device__ __forceinline__ uw_f2 uw_pk(float lo, float hi)
{
    uw_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float uw_lo(uw_f2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float uw_hi(uw_f2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ uw_f2 uw_add2(uw_f2 a, uw_f2 b)
{
    uw_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uw_f2 uw_sub2(uw_f2 a, uw_f2 b)
{
    uw_f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// both products rounded once each; nz must be UW_NEGZERO2 read from a kernel parameter
__device__ __forceinline__ uw_f2 uw_mul2(uw_f2 a, uw_f2 b, uw_f2 nz)
{
    uw_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
// a*b + c with one rounding per lane (callers use it where the product is exact)
__device__ __forceinline__ uw_f2 uw_fma2(uw_f2 a, uw_f2 b, uw_f2 c)
{
    uw_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// scalar times pair (the scalar is broadcast by the instruction's operand form)
__device__ __forceinline__ uw_f2 uw_mul2s(float a, uw_f2 b, uw_f2 nz) { return uw_mul2(uw_pk(a, a), b, nz); }

// launches (defined in the .cu files)
void uw_launch_spectrogram(const UwDims &d, const float2 *x, long long win_stride, int nwin,
                           const float *window, const float2 *twiddle, float *amp, float *ps_dbg,
                           float *psavg, UwPeak *peaks, int *npk, cudaStream_t s);
void uw_launch_worklist(const int *npk, int nwin, int cap, int *base, UwItem *items, int *counters,
                        int *set, cudaStream_t s);
void uw_launch_coarse(const UwDims &d, const float *amp, const UwPeak *peaks, const UwItem *items,
                      const int *total, int cap, const uint32_t *off4, const short *hyp_unique,
                      uwspr_b200_candidate_t *cands, int *ticket, int grid, cudaStream_t s);
// the stage sequence of the fine path over work-list entries *begin + slice0 + [0, slice_n); returns the launches issued
int uw_launch_fine(const UwDims &d, const float2 *x, long long win_stride, const UwItem *items, const int *begin,
                   const int *end, int cap, const uwspr_b200_candidate_t *cands, int jig_first, int jig_count,
                   uwspr_b200_refined_t *refined, uwspr_b200_jiggle_t *jig, uint8_t *soft, int slice0, int slice_n,
                   void *state, void *pbuf, void *pE, void *pbest, void *tables, int *tickets, int grid_points, int grid_lags,
                   int reuse, cudaStream_t s);
size_t uw_fine_state_bytes();   // per candidate of a slice: chain state, stage magnitudes, jiggle magnitudes
size_t uw_fine_pbuf_bytes();
size_t uw_fine_pe_bytes();
size_t uw_fine_pbest_bytes();
size_t uw_fine_tables_bytes();
void uw_fine_blocks_per_sm(int *points, int *lags);
size_t uw_coarse_smem_bytes(const UwDims &d);
int uw_coarse_setup(const UwDims &d);
int uw_fine_setup();
int uw_coarse_blocks_per_sm(const UwDims &d);
