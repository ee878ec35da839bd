// Kernel 2: coarse search over (frequency bin x half-symbol shift x drift hypothesis).
//
// Replaces lib/FDR_impl.cc:339-409 (+ powersum, :188-210) of the reference.  One CTA per
// (window, candidate) taken from the compact work list by an atomic ticket (persistent grid).
//
// Work of one candidate: 5 bins x 26 shifts x (2*maxdrift+1 linear + 125 straight-line
// hypotheses) x 162 symbols.  A hypothesis is a sequence of per-symbol bin offsets; only the
// distinct sequences are evaluated (38 of the 125 trajectories at cf = 1500) and the results
// are looked up when the reference's ordered update rule is replayed.
//
//   stage 1  the candidate's amplitude tile sqrt(ps)[348 rows][if0-5+off_min .. if0+5+off_max]
//            is staged in shared memory (one pass over HBM/L2, 23.7 KB at tile_w 17);
//   stage 2  thread = (shift k0, sequence u) evaluates the five bins if0-2..if0+2 together:
//            per symbol 11 shared loads feed 5 x (ss, pow) chains accumulated in the
//            reference's order (fp32, no FMA, sequential over the 162 symbols);
//   stage 3  warp 0 replays the update rule in scan order (bin, shift, linear drifts, then
//            the 125 trajectories): `sync > cur` for linear, `sync / cur > threshold` for
//            nonlinear, both against the same running `cur` (FDR_impl.cc:360,392).  The
//            rule is an ordered, non-associative fold, so it is replayed, not reduced.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct CoarseLayout {
    int tile_stride;   // floats per tile row (odd, >= tile_w)
    int tile_off, sync_off, off4_off, map_off;  // byte offsets into dynamic smem
    int total;
};

__host__ __device__ inline CoarseLayout coarse_layout(int n_rows, int tile_w, int n_unique, int n_hyp)
{
    CoarseLayout L;
    L.tile_stride = tile_w | 1;
    int o = 0;
    L.tile_off = o;
    o += n_rows * L.tile_stride * 4;
    o = (o + 15) & ~15;
    L.sync_off = o;
    o += UW_NIFR * UW_NK0 * n_unique * 4;
    o = (o + 15) & ~15;
    L.off4_off = o;
    o += UW_NQUAD * n_unique * 4;
    L.map_off = o;
    o += ((n_hyp * 2) + 15) & ~15;
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(kThreads)
k_coarse(UwDims d, const float *__restrict__ amp, const UwPeak *__restrict__ peaks,
         const UwItem *__restrict__ items, const int *__restrict__ total_ptr, int cap,
         const uint32_t *__restrict__ off4_g, const short *__restrict__ hyp_unique_g,
         uwspr_b200_candidate_t *__restrict__ cands, int *__restrict__ ticket)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const CoarseLayout L = coarse_layout(d.n_rows, d.tile_w, d.n_unique, d.n_hyp);
    float *tile = reinterpret_cast<float *>(smem + L.tile_off);
    float *syncv = reinterpret_cast<float *>(smem + L.sync_off);
    uint32_t *off4 = reinterpret_cast<uint32_t *>(smem + L.off4_off);
    short *hmap = reinterpret_cast<short *>(smem + L.map_off);
    __shared__ int s_item;
    __shared__ uint32_t s_sync[6];
    __shared__ float s_gmax[UW_NIFR * UW_NK0];

    const int tid = threadIdx.x, lane = tid & 31;
    const int U = d.n_unique, S = L.tile_stride;
    const int total = min(*total_ptr, cap);

    for (int t = tid; t < UW_NQUAD * U; t += kThreads) off4[t] = off4_g[t];
    for (int t = tid; t < d.n_hyp; t += kThreads) hmap[t] = hyp_unique_g[t];
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 6; q++) s_sync[q] = d.sync_words[q];
    }

    for (;;) {
        __syncthreads();  // previous item fully consumed (and the tables above visible)
        if (tid == 0) s_item = atomicAdd(ticket, 1);
        __syncthreads();
        const int g = s_item;
        if (g >= total) break;
        const UwItem item = items[g];
        const UwPeak pk = peaks[(long long)item.win * d.maxcand + item.slot];
        // FDR_impl.cc:341  if0 = freq/df + m, truncated
        const int if0 = (int)__fadd_rn(__fdiv_rn(pk.freq, d.df), (float)d.m);
        const int col0 = if0 - 5 + d.off_min - d.bin_lo;  // kept-bin index of tile column 0

        // stage 1: amplitude tile
        const float *aw = amp + (long long)item.win * d.n_rows * d.nbp + col0;
        for (int t = tid; t < d.n_rows * d.tile_w; t += kThreads) {
            const int r = t / d.tile_w, c = t - r * d.tile_w;
            tile[r * S + c] = __ldg(aw + (long long)r * d.nbp + c);
        }
        __syncthreads();

        // stage 2: all sums.  task = k0 * U + u (u fastest: the lanes of a warp read the same
        // one or two tile rows, so shared loads are broadcasts / conflict free)
        for (int task = tid; task < UW_NK0 * U; task += kThreads) {
            const int k0 = task / U, u = task - k0 * U;
            // bins 0,1 and 2,3 are carried as packed pairs (FADD2 / FFMA2, common.cuh), bin 4 as scalars: every
            // lane of every instruction performs one of the reference's additions with its own rounding.
            // The sign of the sync bit enters as fma(+-1, d, ss): the product is exact, so the result is the
            // correctly rounded ss +- d, i.e. exactly fadd(ss, +-d).
            // Each symbol reads one copy of the accumulator pairs and writes the other (x -> y -> x ...): ptxas
            // otherwise moves every packed accumulator back to its loop-entry register pair after each symbol.
            uw_f2 ss01 = uw_pk(0.f, 0.f), ss23 = uw_pk(0.f, 0.f), pw01 = uw_pk(0.f, 0.f), pw23 = uw_pk(0.f, 0.f);
            uw_f2 ss01y, ss23y, pw01y, pw23y;
            float ss4 = 0.f, pw4 = 0.f;
            const float *rowp = tile + k0 * S;
#define UW_COARSE_SYMBOL(E, SS01I, SS23I, PW01I, PW23I, SS01O, SS23O, PW01O, PW23O)                                 \
    do {                                                                                                            \
        const float *A = rowp + ((w >> (8 * (E))) & 0xffu);                                                         \
        const uw_f2 a01 = uw_pk(A[0], A[1]), a23 = uw_pk(A[2], A[3]), a45 = uw_pk(A[4], A[5]);                      \
        const uw_f2 a67 = uw_pk(A[6], A[7]), a89 = uw_pk(A[8], A[9]);                                               \
        const float a10 = A[10];                                                                                    \
        /* pair sums A[x] + A[x+4]: (p0+p2) of bin x and (p1+p3) of bin x-2 */                                      \
        const uw_f2 s01 = uw_add2(a01, a45), s23 = uw_add2(a23, a67), s45 = uw_add2(a45, a89);                      \
        const float s6 = __fadd_rn(uw_lo(a67), a10);                                                                \
        /* powersum(): ss += (2*pr3[k]-1) * ((p1+p3)-(p0+p2)) */                                                    \
        const float sgn = ((sbits >> (E)) & 1u) ? 1.0f : -1.0f;                                                     \
        SS01O = uw_fma2(uw_pk(sgn, sgn), uw_sub2(s23, s01), SS01I);                                                 \
        SS23O = uw_fma2(uw_pk(sgn, sgn), uw_sub2(s45, s23), SS23I);                                                 \
        ss4 = fmaf(sgn, __fsub_rn(s6, uw_lo(s45)), ss4);                                                            \
        /* pow = pow + p0 + p1 + p2 + p3, left to right */                                                          \
        PW01O = uw_add2(uw_add2(uw_add2(uw_add2(PW01I, a01), a23), a45), a67);                                      \
        PW23O = uw_add2(uw_add2(uw_add2(uw_add2(PW23I, a23), a45), a67), a89);                                      \
        pw4 = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(pw4, uw_lo(a45)), uw_lo(a67)), uw_lo(a89)), a10);             \
        rowp += 2 * S;                                                                                              \
    } while (0)
            for (int q = 0; q < UW_NQUAD; q++) {
                const uint32_t w = off4[q * U + u];
                const uint32_t sbits = s_sync[(4 * q) >> 5] >> ((4 * q) & 31);  // 4 | 32: no straddle
                UW_COARSE_SYMBOL(0, ss01, ss23, pw01, pw23, ss01y, ss23y, pw01y, pw23y);
                UW_COARSE_SYMBOL(1, ss01y, ss23y, pw01y, pw23y, ss01, ss23, pw01, pw23);
                static_assert(UW_NSYM % 4 == 2, "the last quad holds two symbols");
                if (4 * q + 2 < UW_NSYM) {
                    UW_COARSE_SYMBOL(2, ss01, ss23, pw01, pw23, ss01y, ss23y, pw01y, pw23y);
                    UW_COARSE_SYMBOL(3, ss01y, ss23y, pw01y, pw23y, ss01, ss23, pw01, pw23);
                }
            }
#undef UW_COARSE_SYMBOL
            const float ss0 = uw_lo(ss01), ss1 = uw_hi(ss01), ss2 = uw_lo(ss23), ss3 = uw_hi(ss23);
            const float pw0 = uw_lo(pw01), pw1 = uw_hi(pw01), pw2 = uw_lo(pw23), pw3 = uw_hi(pw23);
            float *sv = syncv + k0 * U + u;
            sv[0 * UW_NK0 * U] = __fdiv_rn(ss0, pw0);
            sv[1 * UW_NK0 * U] = __fdiv_rn(ss1, pw1);
            sv[2 * UW_NK0 * U] = __fdiv_rn(ss2, pw2);
            sv[3 * UW_NK0 * U] = __fdiv_rn(ss3, pw3);
            sv[4 * UW_NK0 * U] = __fdiv_rn(ss4, pw4);
        }
        __syncthreads();
        // largest sum of every (bin, shift) group (NaNs ignored), for the skip test of stage 3
        if (tid < UW_NIFR * UW_NK0) {
            const float *sv = syncv + tid * U;
            float mx = -INFINITY;
            for (int u = 0; u < U; u++) mx = fmaxf(mx, sv[u]);
            s_gmax[tid] = mx;
        }
        __syncthreads();

        // stage 3: ordered replay of the update rule by warp 0
        if (tid < 32) {
            float cur = -1e30f;  // :340
            int b_type = -1, b_idx = 0, b_k0 = 0, b_a = 0;
            // Once cur is positive and threshold >= 1, a group whose largest sum does not exceed cur
            // cannot change anything: the linear rule needs v > cur, and v <= cur gives a correctly
            // rounded v / cur <= 1 <= threshold.  Most of the 130 groups are skipped this way.
            const bool can_skip = d.threshold >= 1.0f;
            for (int a = 0; a < UW_NIFR; a++) {
                for (int k0 = 0; k0 < UW_NK0; k0++) {
                    if (can_skip && cur > 0.0f && !(s_gmax[a * UW_NK0 + k0] > cur)) continue;
                    const float *sv = syncv + (a * UW_NK0 + k0) * U;
                    // linear drifts in order: running strict maximum == first maximum above cur
                    for (int base = 0; base < d.n_lin; base += 32) {
                        const int h = base + lane;
                        const float v = (h < d.n_lin) ? sv[hmap[h]] : 0.0f;
                        const bool ok = (h < d.n_lin) && (v > cur);
                        const unsigned any = __ballot_sync(0xffffffffu, ok);
                        if (any) {
                            const float mx = warp_max(ok ? v : -INFINITY);
                            const unsigned at = __ballot_sync(0xffffffffu, ok && v == mx);
                            cur = mx;
                            b_type = 0;
                            b_idx = base + __ffs(at) - 1;
                            b_k0 = k0;
                            b_a = a;
                        }
                    }
                    // trajectories in generator order: each hit changes cur, so rescan from the hit
                    int pos = 0;
                    for (;;) {
                        int hit = -1;
                        for (int base = 0; base < UW_NTRAJ && hit < 0; base += 32) {
                            const int h = base + lane;
                            bool ok = false;
                            if (h < UW_NTRAJ && h >= pos) ok = __fdiv_rn(sv[hmap[d.n_lin + h]], cur) > d.threshold;
                            const unsigned b = __ballot_sync(0xffffffffu, ok);
                            if (b) hit = base + __ffs(b) - 1;
                        }
                        if (hit < 0) break;
                        cur = sv[hmap[d.n_lin + hit]];
                        b_type = 1;
                        b_idx = hit;
                        b_k0 = k0;
                        b_a = a;
                        pos = hit + 1;
                    }
                }
            }
            if (lane == 0) {
                uwspr_b200_candidate_t c;
                unsigned long long *raw = reinterpret_cast<unsigned long long *>(&c);
#pragma unroll
                for (int q = 0; q < 6; q++) raw[q] = 0ull;
                c.snr = pk.snr;
                c.sync = cur;
                if (b_type < 0) {
                    c.freq = pk.freq;  // never updated (all sums NaN): the reference keeps the peak frequency
                } else {
                    c.freq = __fmul_rn((float)(if0 - 2 + b_a - d.m), d.df);  // :362
                    c.shift = 128 * b_k0;                                      // :361
                    if (b_type == 0) {
                        c.m_type = 0;
                        c.m_linear.drift = (float)(b_idx - d.maxdrift);
                    } else {
                        c.m_type = 1;  // slm.cc:76-116: p2 index fastest, then V1, then V2
                        c.m_nonlinear.V1 = (double)((b_idx / 5) % 5) - 2.0;
                        c.m_nonlinear.V2 = (double)(b_idx / 25) - 2.0;
                        c.m_nonlinear.p1 = 0;
                        c.m_nonlinear.p2 = (b_idx % 5) * 200 + 50;
                    }
                }
                cands[g] = c;
            }
        }
    }
}

}  // namespace

size_t uw_coarse_smem_bytes(const UwDims &d)
{
    return (size_t)coarse_layout(d.n_rows, d.tile_w, d.n_unique, d.n_hyp).total;
}

// The attribute is per function, not per context: it is raised to the device's opt-in
// maximum once, so contexts with different table sizes never lower each other's limit.
int uw_coarse_setup(const UwDims &d)
{
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 1;
    if ((size_t)optin < uw_coarse_smem_bytes(d)) return 1;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k_coarse) != cudaSuccess) return 1;
    cudaError_t e = cudaFuncSetAttribute(k_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         optin - (int)fa.sharedSizeBytes);
    return e == cudaSuccess ? 0 : 1;
}

void uw_launch_coarse(const UwDims &d, const float *amp, const UwPeak *peaks, const UwItem *items,
                      const int *total, int cap, const uint32_t *off4, const short *hyp_unique,
                      uwspr_b200_candidate_t *cands, int *ticket, int grid, cudaStream_t s)
{
    k_coarse<<<grid, kThreads, uw_coarse_smem_bytes(d), s>>>(d, amp, peaks, items, total, cap, off4,
                                                             hyp_unique, cands, ticket);
}

int uw_coarse_blocks_per_sm(const UwDims &d)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_coarse, kThreads, uw_coarse_smem_bytes(d)) != cudaSuccess) return 1;
    return n > 0 ? n : 1;
}
