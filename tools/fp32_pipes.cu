// Micro-benchmark of the non-fused fp32 pipes on sm_100a: scalar add/mul vs the packed
// add.rn.f32x2 / mul.rn.f32x2 forms.  Prints lane-operations per clock per SM for each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_pipes fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kIlp = 8;

__device__ __forceinline__ unsigned long long pack(float a, float b)
{
    return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

template <int MODE> __global__ void k(float *out, float seed)
{
    float a[kIlp], b = seed;
    unsigned long long p[kIlp], q = pack(seed, seed * 0.5f);
#pragma unroll
    for (int i = 0; i < kIlp; i++) {
        a[i] = seed + i + threadIdx.x;
        p[i] = pack(a[i], a[i] * 2.0f);
    }
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kIlp; i++) {
            if (MODE == 0) a[i] = __fadd_rn(a[i], b);
            if (MODE == 1) a[i] = __fmul_rn(a[i], b);
            if (MODE == 2) a[i] = __fadd_rn(__fmul_rn(a[i], b), b);          // mul then add, two instructions
            if (MODE == 3) p[i] = add2(p[i], q);
            if (MODE == 4) p[i] = mul2(p[i], q);
            if (MODE == 5) p[i] = add2(mul2(p[i], q), q);
            if (MODE == 6) a[i] = fmaf(a[i], b, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kIlp; i++) s += a[i] + __uint_as_float((unsigned)(p[i] & 0xffffffffu)) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, double lane_ops_per_iter, float *out, int sms, double mhz)
{
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * kIters * kIlp * lane_ops_per_iter;
    printf("%-28s %8.3f ms  %7.2f Tlane-op/s  %6.1f lane-ops/clk/SM (at %.0f MHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    float *out;
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
    printf("%s, %d SMs, max clock %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    run<0>("fadd scalar", 1, out, p.multiProcessorCount, mhz);
    run<1>("fmul scalar", 1, out, p.multiProcessorCount, mhz);
    run<2>("fmul+fadd scalar", 2, out, p.multiProcessorCount, mhz);
    run<3>("add.f32x2", 2, out, p.multiProcessorCount, mhz);
    run<4>("mul.f32x2", 2, out, p.multiProcessorCount, mhz);
    run<5>("mul.f32x2 + add.f32x2", 4, out, p.multiProcessorCount, mhz);
    run<6>("ffma scalar (1 lane-op)", 1, out, p.multiProcessorCount, mhz);
    return 0;
}
