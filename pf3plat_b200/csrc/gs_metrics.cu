// gs_metrics.cu -- PSNR of rendered views against ground truth.
//
// Semantics: compute_psnr of /root/reference/src/evaluation/metrics.py:11-19 -- clip both images to [0,1], mean
// squared error over (c,h,w) per image, -10 log10(mse).  One streaming pass with 16-byte loads, per-CTA partial
// sums, a deterministic fp64 finalisation (no atomics, so the result is bit-reproducible).
#include "gs_common.cuh"

namespace {

constexpr int PS_THREADS = 256;

__device__ __forceinline__ float clip01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

__global__ void __launch_bounds__(PS_THREADS)
k_sq_err_partial(const float *__restrict__ gt, const float *__restrict__ pred, int64_t n, float *__restrict__ partial) {
    const int img = blockIdx.y;
    const float *a = gt + (size_t)img * n, *b = pred + (size_t)img * n;
    float acc = 0.f;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
    const int64_t n4 = vec ? n / 4 : 0;
    for (int64_t i = (int64_t)blockIdx.x * PS_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * PS_THREADS) {
        const float4 x = reinterpret_cast<const float4 *>(a)[i], y = reinterpret_cast<const float4 *>(b)[i];
        const float d0 = clip01(x.x) - clip01(y.x), d1 = clip01(x.y) - clip01(y.y);
        const float d2 = clip01(x.z) - clip01(y.z), d3 = clip01(x.w) - clip01(y.w);
        acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * PS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PS_THREADS) {
        const float d = clip01(a[i]) - clip01(b[i]);
        acc += d * d;
    }
    __shared__ float s[PS_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < PS_THREADS / 32; w++) t += s[w];
        partial[(size_t)img * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void k_psnr_finalize(const float *__restrict__ partial, int nblocks, int64_t n, float *__restrict__ out) {
    const int img = blockIdx.x;
    double t = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += 32) t += (double)partial[(size_t)img * nblocks + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) out[img] = (float)(-10.0 * log10(t / (double)n));
}

}  // namespace

// scratch: at least gs_psnr_scratch_floats(batch, n) floats
extern "C" GS_API int64_t gs_psnr_scratch_floats(int32_t batch, int64_t n) {
    int64_t blocks = (n / 4 + PS_THREADS - 1) / PS_THREADS;
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    return (int64_t)batch * blocks;
}

extern "C" GS_API int gs_psnr(const float *ground_truth, const float *predicted, int32_t batch, int64_t n, float *scratch,
                              float *out, void *stream) {
    if (!ground_truth || !predicted || !scratch || !out || batch < 0 || n < 1) return gs_set_error(GS_ERR_INVALID, "bad gs_psnr arguments");
    if (batch == 0) return GS_OK;
    const int blocks = (int)(gs_psnr_scratch_floats(batch, n) / batch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    k_sq_err_partial<<<dim3(blocks, batch), PS_THREADS, 0, st>>>(ground_truth, predicted, n, scratch);
    GS_CUDA_OK(cudaGetLastError());
    k_psnr_finalize<<<batch, 32, 0, st>>>(scratch, blocks, n, out);
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}
