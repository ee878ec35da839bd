// Micro-benchmark of the non-fused fp32 paths on sm_100a.
//
// The reference's sums are decided by separately rounded fp32 products and additions, so the
// kernels cannot use FMA contraction.  This tool measures what the SM sustains for
//   * scalar FADD / FMUL / FMUL+FADD
//   * packed add.rn.f32x2 (FADD2) and the packed product issued as fma.rn.f32x2(a, b, -0.0)
//     (FFMA2) with the -0.0 pair in a register, in a uniform register (direct kernel
//     parameter) or in constant memory
//   * the inner loop of k_fine (two correlations + table rotation per tone and sample) in its
//     scalar and packed forms, operands in registers
// and prints lane-operations per clock per SM (one lane-op = one fp32 add or multiply of one
// thread).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_pipes fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
constexpr int kIters = 2048;
constexpr int kIlp = 8;

__constant__ u64 g_nz = 0x8000000080000000ull;

__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b, u64 nz) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz)); return r; }

enum { FADD, FMUL, FMULADD, FFMA, ADD2, MUL2_REG, MUL2_UR, MUL2_CONST, LOOP_SCALAR, LOOP_PACKED_REG, LOOP_PACKED_UR, NMODES };

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float seed, u64 nz_param, const u64 *nz_mem)
{
    float a[kIlp], b = seed;
    u64 p[kIlp], q = pack(seed, seed * 0.5f);
    u64 nz = nz_param;
    if (MODE == MUL2_REG || MODE == LOOP_PACKED_REG) nz = nz_mem[threadIdx.x & 1];   // per-thread load: lives in registers
    if (MODE == MUL2_CONST) nz = g_nz;
#pragma unroll
    for (int i = 0; i < kIlp; i++) {
        a[i] = seed + i + threadIdx.x;
        p[i] = pack(a[i], a[i] * 2.0f);
    }
    if (MODE < LOOP_SCALAR) {
        for (int it = 0; it < kIters; it++) {
#pragma unroll
            for (int i = 0; i < kIlp; i++) {
                if (MODE == FADD) a[i] = __fadd_rn(a[i], b);
                if (MODE == FMUL) a[i] = __fmul_rn(a[i], b);
                if (MODE == FMULADD) a[i] = __fadd_rn(__fmul_rn(a[i], b), b);
                if (MODE == FFMA) a[i] = fmaf(a[i], b, b);
                if (MODE == ADD2) p[i] = add2(p[i], q);
                if (MODE == MUL2_REG || MODE == MUL2_UR || MODE == MUL2_CONST) p[i] = mul2(p[i], q, nz);
            }
        }
    } else if (MODE == LOOP_SCALAR) {
        // four tones: c, s, cd, sd, inp, quad per tone; x varies per iteration
        float c[4], s[4], cd[4], sd[4], inp[4], quad[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { c[j] = 1.f; s[j] = 0.f; cd[j] = cosf(seed * (j + 1)); sd[j] = sinf(seed * (j + 1)); inp[j] = 0.f; quad[j] = 0.f; }
        float xr = seed, xi = seed * 0.5f;
        for (int it = 0; it < kIters; it++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                inp[j] = __fadd_rn(__fadd_rn(inp[j], __fmul_rn(xr, c[j])), __fmul_rn(xi, s[j]));
                quad[j] = __fadd_rn(__fsub_rn(quad[j], __fmul_rn(xr, s[j])), __fmul_rn(xi, c[j]));
                const float cn = __fsub_rn(__fmul_rn(c[j], cd[j]), __fmul_rn(s[j], sd[j]));
                const float sn = __fadd_rn(__fmul_rn(c[j], sd[j]), __fmul_rn(s[j], cd[j]));
                c[j] = cn; s[j] = sn;
            }
            xr = __int_as_float(__float_as_int(xr) ^ (it & 1));   // alu-pipe work standing in for the loads
            xi = __int_as_float(__float_as_int(xi) ^ (it & 2));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = inp[j] + quad[j];
    } else {
        u64 c[2], s[2], cd[2], sd[2], inp[2], quad[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            c[j] = pack(1.f, 1.f); s[j] = pack(0.f, 0.f);
            cd[j] = pack(cosf(seed * (2 * j + 1)), cosf(seed * (2 * j + 2)));
            sd[j] = pack(sinf(seed * (2 * j + 1)), sinf(seed * (2 * j + 2)));
            inp[j] = pack(0.f, 0.f); quad[j] = pack(0.f, 0.f);
        }
        float xr = seed, xi = seed * 0.5f;
        for (int it = 0; it < kIters; it++) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                inp[j] = add2(add2(inp[j], mul2(pack(xr, xr), c[j], nz)), mul2(pack(xi, xi), s[j], nz));
                quad[j] = add2(sub2(quad[j], mul2(pack(xr, xr), s[j], nz)), mul2(pack(xi, xi), c[j], nz));
                const u64 cn = sub2(mul2(c[j], cd[j], nz), mul2(s[j], sd[j], nz));
                const u64 sn = add2(mul2(c[j], sd[j], nz), mul2(s[j], cd[j], nz));
                c[j] = cn; s[j] = sn;
            }
            xr = __int_as_float(__float_as_int(xr) ^ (it & 1));
            xi = __int_as_float(__float_as_int(xi) ^ (it & 2));
        }
#pragma unroll
        for (int j = 0; j < 2; j++) p[j] = add2(inp[j], quad[j]);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < kIlp; i++) r += a[i] + lo(p[i]) + hi(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char *name, double lane_ops_per_thread, float *out, const u64 *nz_mem, int sms, double mhz, int ctas_per_sm)
{
    const int blocks = sms * ctas_per_sm, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0x8000000080000000ull, nz_mem);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0x8000000080000000ull, nz_mem);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * lane_ops_per_thread;
    printf("%-44s %d CTA/SM %8.3f ms  %7.2f Tlane-op/s  %6.1f lane-ops/clk/SM (at %.0f MHz)\n", name, ctas_per_sm, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    const int sms = p.multiProcessorCount;
    float *out;
    u64 *nz_mem;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    cudaMalloc(&nz_mem, 2 * sizeof(u64));
    const u64 nzh[2] = { 0x8000000080000000ull, 0x8000000080000000ull };
    cudaMemcpy(nz_mem, nzh, sizeof(nzh), cudaMemcpyHostToDevice);
    printf("%s, %d SMs, max clock %.0f MHz\n", p.name, sms, mhz);
    const double simple = (double)kIters * kIlp, loop = (double)kIters * 4 * 14;
    for (int ctas : { 8, 3 }) {
        run<FADD>("fadd scalar", simple, out, nz_mem, sms, mhz, ctas);
        run<FMUL>("fmul scalar", simple, out, nz_mem, sms, mhz, ctas);
        run<FMULADD>("fmul+fadd scalar (2 instr)", 2 * simple, out, nz_mem, sms, mhz, ctas);
        run<FFMA>("ffma scalar (counted as 1 lane-op)", simple, out, nz_mem, sms, mhz, ctas);
        run<ADD2>("add.f32x2", 2 * simple, out, nz_mem, sms, mhz, ctas);
        run<MUL2_REG>("fma.f32x2(a,b,-0) -0 in registers", 2 * simple, out, nz_mem, sms, mhz, ctas);
        run<MUL2_UR>("fma.f32x2(a,b,-0) -0 in a uniform register", 2 * simple, out, nz_mem, sms, mhz, ctas);
        run<MUL2_CONST>("fma.f32x2(a,b,-0) -0 in constant memory", 2 * simple, out, nz_mem, sms, mhz, ctas);
        run<LOOP_SCALAR>("k_fine inner loop, scalar (56 instr/sample)", loop, out, nz_mem, sms, mhz, ctas);
        run<LOOP_PACKED_REG>("k_fine inner loop, packed, -0 in registers", loop, out, nz_mem, sms, mhz, ctas);
        run<LOOP_PACKED_UR>("k_fine inner loop, packed, -0 uniform", loop, out, nz_mem, sms, mhz, ctas);
    }
    return 0;
}
