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

// launches (defined in the .cu files)
void uw_launch_spectrogram(const UwDims &d, const float2 *x, long long win_stride, int nwin,
                           const float *window, const float2 *twiddle, float *amp, float *ps_dbg,
                           float *psavg, UwPeak *peaks, int *npk, cudaStream_t s);
void uw_launch_worklist(const int *npk, int nwin, int cap, int *base, UwItem *items, int *counters,
                        int *set, cudaStream_t s);
void uw_launch_coarse(const UwDims &d, const float *amp, const UwPeak *peaks, const UwItem *items,
                      const int *total, int cap, const uint32_t *off4, const short *hyp_unique,
                      uwspr_b200_candidate_t *cands, int *ticket, int grid, cudaStream_t s);
void uw_launch_fine(const UwDims &d, const float2 *x, long long win_stride, const UwItem *items,
                    const int *total, int cap, const uwspr_b200_candidate_t *cands, int jig_first,
                    int jig_count, uwspr_b200_refined_t *refined, uwspr_b200_jiggle_t *jig,
                    uint8_t *soft, int *ticket, int grid, cudaStream_t s);
size_t uw_coarse_smem_bytes(const UwDims &d);
size_t uw_fine_smem_bytes();
int uw_coarse_setup(const UwDims &d);
int uw_fine_setup();
int uw_coarse_blocks_per_sm(const UwDims &d);
int uw_fine_blocks_per_sm();
