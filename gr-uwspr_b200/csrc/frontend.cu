// Front-end ahead of the hot path: real audio at fs_in (12 kHz in the
// reference flowgraphs) -> mix down by fc -> FIR low-pass -> keep every `decim`-th sample ->
// complex64 at 375 sps, the stream uwspr.sliding_window_stream_to_pdu consumes.
//
// In the reference this is done by stock GNU Radio blocks, not by gr-uwspr code
// (examples/WaveFilePlusNoiseDecode.grc:834-958 two freq_xlating_fft_filter_ccc, :1753-1810
// rational_resampler decim 32).  One translating decimating FIR with caller-supplied taps
// replaces the cascade:
//
//     y[m] = sum_k taps[k] * x[n - k] * exp(-2 pi i fc (n - k) / fs_in),  n = m * decim + delay
//
// with x[n] = 0 outside [0, n_in) (the zero history of a GNU Radio FIR).  It pays off when many
// channels of audio are already on the device (the 64-hydrophone array of BASELINE.json
// configs[4]): the 375-sps output feeds uwspr_b200_coarse_fine as a device pointer.
// taps may be complex (uwspr_b200_frontend_ctaps): the flowgraph's cascade - real band-pass around fc,
// translate by fc + low-pass, rational resampler 1/32 - collapses into one complex composite filter
// (h_bp[k] e^{-i w k}) * h_lp * h_rs applied to the mixed-down input (uwspr_b200/binding.py flowgraph_taps).
// GNU Radio itself is not available here: oracle/gr_frontend.py restates its blocks one by one in
// float64 from their published algorithms (gr-filter 3.7: firdes, freq_xlating_fft_filter, rational_resampler)
// and the GPU test compares this kernel with that chain; tests/test_gpu_parity.py also checks the formula above
// against a float64 numpy evaluation.
#include <string>

#include "common.cuh"

namespace {

constexpr int kFeOut = 64;      // outputs per CTA
constexpr int kFeThreads = 256; // 8 warps x 8 outputs

template <bool CTAPS>
__global__ void __launch_bounds__(kFeThreads)
k_frontend(const void *__restrict__ audio, int fmt, long long chan_stride, long long n_in,
           const float *__restrict__ taps, int ntaps, int decim, int delay, double cyc_per_sample,
           float2 *__restrict__ out, long long out_stride, long long n_out)
{
    extern __shared__ __align__(16) unsigned char fe_smem[];
    float2 *z = reinterpret_cast<float2 *>(fe_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long m0 = (long long)blockIdx.x * kFeOut;
    const int chan = blockIdx.y;
    const int span = (kFeOut - 1) * decim + ntaps;
    const long long nb = m0 * decim + delay - (ntaps - 1);   // input index of z[0]
    const float *af = reinterpret_cast<const float *>(audio) + (long long)chan * chan_stride;
    const short *as = reinterpret_cast<const short *>(audio) + (long long)chan * chan_stride;
    for (int t = tid; t < span; t += kFeThreads) {
        const long long n = nb + t;
        float v = 0.0f;
        if (n >= 0 && n < n_in) v = fmt ? (float)as[n] * (1.0f / 32768.0f) : af[n];
        // local oscillator exp(-2 pi i fc n / fs): the phase in turns, reduced in double
        double turns = (double)n * cyc_per_sample;
        turns -= floor(turns);
        double sn, cs;
        sincospi(2.0 * turns, &sn, &cs);
        z[t] = make_float2(v * (float)cs, -v * (float)sn);
    }
    __syncthreads();
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; o++) acc[o] = make_float2(0.0f, 0.0f);
    const float2 *zw = z + (warp * 8) * decim + (ntaps - 1);
    for (int k = lane; k < ntaps; k += 32) {
        const float h = CTAPS ? __ldg(taps + 2 * k) : __ldg(taps + k);
        const float hi = CTAPS ? __ldg(taps + 2 * k + 1) : 0.0f;
#pragma unroll
        for (int o = 0; o < 8; o++) {
            const float2 s = zw[o * decim - k];
            acc[o].x = fmaf(h, s.x, acc[o].x);
            acc[o].y = fmaf(h, s.y, acc[o].y);
            if (CTAPS) {
                acc[o].x = fmaf(-hi, s.y, acc[o].x);
                acc[o].y = fmaf(hi, s.x, acc[o].y);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 8; o++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            acc[o].x += __shfl_xor_sync(0xffffffffu, acc[o].x, d);
            acc[o].y += __shfl_xor_sync(0xffffffffu, acc[o].y, d);
        }
        const long long m = m0 + warp * 8 + o;
        if (lane == 0 && m < n_out) out[(long long)chan * out_stride + m] = acc[o];
    }
}

thread_local std::string g_fe_error;

int fe_fail(int st, const std::string &msg)
{
    g_fe_error = msg;
    return st;
}

}  // namespace

#define FE_CU(call)                                                                                \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            if (d_audio_own) cudaFree(d_audio_own);                                                \
            if (d_out_own) cudaFree(d_out_own);                                                    \
            if (d_taps) cudaFree(d_taps);                                                          \
            return fe_fail(UWSPR_B200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
        }                                                                                          \
    } while (0)

extern "C" const char *uwspr_b200_frontend_error(void) { return g_fe_error.c_str(); }

static int frontend_impl(bool ctaps, int device, const void *audio, int fmt, int space_in, int64_t chan_stride, int nchan,
                         int64_t n_in, const float *taps, int ntaps, int decim, int delay, double fc,
                         double fs_in, float *out, int space_out, int64_t out_stride, int64_t *n_out_p)
{
    if (!audio || !taps || !out || nchan < 1 || n_in < 1 || ntaps < 1 || ntaps > 16384 || decim < 1 || decim > 4096 ||
        delay < 0 || !(fs_in > 0.0) || (fmt != 0 && fmt != 1) || chan_stride < n_in)
        return fe_fail(UWSPR_B200_E_PARAM, "bad front-end arguments");
    if (nchan > 65535) return fe_fail(UWSPR_B200_E_PARAM, "more than 65535 channels in one call (grid.y limit)");
    const int64_t n_out = n_in / decim;
    if (n_out_p) *n_out_p = n_out;
    if (n_out == 0) return UWSPR_B200_OK;
    if (out_stride < n_out) return fe_fail(UWSPR_B200_E_PARAM, "out_stride shorter than the output");
    void *d_audio_own = nullptr, *d_out_own = nullptr;
    float *d_taps = nullptr;
    FE_CU(cudaSetDevice(device));
    const size_t esz = fmt ? sizeof(short) : sizeof(float);
    const void *d_audio = audio;
    if (space_in == UWSPR_B200_HOST) {
        const size_t bytes = ((size_t)(nchan - 1) * (size_t)chan_stride + (size_t)n_in) * esz;
        FE_CU(cudaMalloc(&d_audio_own, bytes));
        FE_CU(cudaMemcpy(d_audio_own, audio, bytes, cudaMemcpyHostToDevice));
        d_audio = d_audio_own;
    }
    float2 *d_out = reinterpret_cast<float2 *>(out);
    const size_t out_bytes = ((size_t)(nchan - 1) * (size_t)out_stride + (size_t)n_out) * sizeof(float2);
    if (space_out == UWSPR_B200_HOST) {
        FE_CU(cudaMalloc(&d_out_own, out_bytes));
        d_out = reinterpret_cast<float2 *>(d_out_own);
    }
    const size_t tap_bytes = sizeof(float) * (size_t)ntaps * (ctaps ? 2 : 1);
    FE_CU(cudaMalloc(&d_taps, tap_bytes));
    FE_CU(cudaMemcpy(d_taps, taps, tap_bytes, cudaMemcpyHostToDevice));
    const size_t smem = ((size_t)(kFeOut - 1) * decim + ntaps) * sizeof(float2);
    int optin = 0;
    FE_CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem > (size_t)optin) {
        cudaFree(d_taps);
        if (d_audio_own) cudaFree(d_audio_own);
        if (d_out_own) cudaFree(d_out_own);
        return fe_fail(UWSPR_B200_E_PARAM, "ntaps + 63*decim samples do not fit shared memory");
    }
    FE_CU(cudaFuncSetAttribute(k_frontend<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    FE_CU(cudaFuncSetAttribute(k_frontend<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    const dim3 grid((unsigned)((n_out + kFeOut - 1) / kFeOut), (unsigned)nchan);
    if (ctaps)
        k_frontend<true><<<grid, kFeThreads, smem>>>(d_audio, fmt, (long long)chan_stride, (long long)n_in, d_taps, ntaps, decim,
                                                     delay, fc / fs_in, d_out, (long long)out_stride, (long long)n_out);
    else
        k_frontend<false><<<grid, kFeThreads, smem>>>(d_audio, fmt, (long long)chan_stride, (long long)n_in, d_taps, ntaps, decim,
                                                      delay, fc / fs_in, d_out, (long long)out_stride, (long long)n_out);
    FE_CU(cudaGetLastError());
    if (space_out == UWSPR_B200_HOST)
        // channel by channel: the caller's memory between two channels (out_stride > n_out) is not touched
        FE_CU(cudaMemcpy2D(out, (size_t)out_stride * sizeof(float2), d_out_own, (size_t)out_stride * sizeof(float2),
                           (size_t)n_out * sizeof(float2), (size_t)nchan, cudaMemcpyDeviceToHost));
    else
        FE_CU(cudaDeviceSynchronize());
    cudaFree(d_taps);
    if (d_audio_own) cudaFree(d_audio_own);
    if (d_out_own) cudaFree(d_out_own);
    return UWSPR_B200_OK;
}

extern "C" int uwspr_b200_frontend(int device, const void *audio, int fmt, int space_in, int64_t chan_stride, int nchan,
                                   int64_t n_in, const float *taps, int ntaps, int decim, int delay, double fc,
                                   double fs_in, float *out, int space_out, int64_t out_stride, int64_t *n_out_p)
{
    return frontend_impl(false, device, audio, fmt, space_in, chan_stride, nchan, n_in, taps, ntaps, decim, delay, fc, fs_in, out,
                         space_out, out_stride, n_out_p);
}

extern "C" int uwspr_b200_frontend_ctaps(int device, const void *audio, int fmt, int space_in, int64_t chan_stride, int nchan,
                                         int64_t n_in, const float *taps_iq, int ntaps, int decim, int delay, double fc,
                                         double fs_in, float *out, int space_out, int64_t out_stride, int64_t *n_out_p)
{
    return frontend_impl(true, device, audio, fmt, space_in, chan_stride, nchan, n_in, taps_iq, ntaps, decim, delay, fc, fs_in, out,
                         space_out, out_stride, n_out_p);
}
