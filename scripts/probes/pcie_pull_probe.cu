// pcie_pull_probe.cu -- how fast can a kernel PULL only the first `take` bytes of every `row` bytes out of pinned host
// memory (zero-copy over PCIe), compared with the copy engine moving whole rows?  Question behind it: gs_render_host
// uploads 300-byte SH rows of which the evaluator reads 192 (bands 0..3); cudaMemcpy2DAsync with 192-byte rows was
// measured far below PCIe rate in round 1.   Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pcie_pull_probe pcie_pull_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// one warp per row group: thread t copies float t of the wanted part of its row (coalesced 4-byte loads)
__global__ void pull_rows_f32(const float *__restrict__ src, float *__restrict__ dst, int rows, int row_f, int take_f) {
    const long long total = (long long)rows * take_f;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / take_f), c = (int)(e - (long long)r * take_f);
        dst[e] = __ldg(src + (long long)r * row_f + c);
    }
}
// 16-byte loads where the source address allows it: rows start 16-byte aligned every 4th row (300 = 16*18 + 12)
__global__ void pull_rows_f32x4(const float *__restrict__ src, float *__restrict__ dst, int rows, int row_f, int take_f) {
    // 4 rows = 1200 bytes = 75 float4: handle groups of 4 rows; wanted floats of the group: rows k*75 .. k*75+take
    const int groups = rows / 4;
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int g = warp; g < groups; g += nwarps) {
        const float4 *s4 = reinterpret_cast<const float4 *>(src + (long long)g * 4 * row_f);
        for (int q = lane; q < row_f; q += 32) {  // row_f float4 per group of 4 rows
            // float4 q covers floats [4q, 4q+4) of the group; keep it if any of them is wanted
            const int f0 = 4 * q, r0 = f0 / row_f, c0 = f0 - r0 * row_f, f3 = f0 + 3, r3 = f3 / row_f, c3 = f3 - r3 * row_f;
            if (c0 < take_f || c3 < take_f || r3 != r0) {
                const float4 v = __ldg(s4 + q);
                const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int f = f0 + k, r = f / row_f, c = f - r * row_f;
                    if (c < take_f) dst[((long long)g * 4 + r) * take_f + c] = vals[k];
                }
            }
        }
    }
}
__global__ void pull_all_f32x4(const float4 *__restrict__ src, float4 *__restrict__ dst, long long n4) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) dst[e] = __ldg(src + e);
}

int main() {
    const int rows = 500000, row_f = 75, take_f = 48;
    const size_t src_bytes = (size_t)rows * row_f * 4, dst_bytes = (size_t)rows * take_f * 4;
    float *h, *d_full, *d_take;
    CK(cudaHostAlloc(&h, src_bytes, cudaHostAllocDefault));
    for (size_t i = 0; i < src_bytes / 4; i++) h[i] = (float)(i & 1023);
    CK(cudaMalloc(&d_full, src_bytes));
    CK(cudaMalloc(&d_take, dst_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](const char *name, size_t useful, auto fn) {
        fn(); CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int it = 0; it < 5; it++) {
            CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("%-44s %8.3f ms  useful %6.1f MB -> %6.1f GB/s useful, %6.1f GB/s of the whole block\n", name, best, useful / 1e6,
               useful / best / 1e6, src_bytes / best / 1e6);
    };
    time("cudaMemcpyAsync whole rows (150 MB)", src_bytes, [&] { CK(cudaMemcpyAsync(d_full, h, src_bytes, cudaMemcpyHostToDevice)); });
    time("cudaMemcpy2DAsync 192 of 300 B", dst_bytes, [&] { CK(cudaMemcpy2DAsync(d_take, take_f * 4, h, row_f * 4, take_f * 4, rows, cudaMemcpyHostToDevice)); });
    for (int ctas : {148, 296, 592, 1184}) {
        char name[96];
        snprintf(name, sizeof(name), "kernel pull all, float4, %d CTAs x 256", ctas);
        time(name, src_bytes, [&] { pull_all_f32x4<<<ctas, 256>>>((const float4 *)h, (float4 *)d_full, (long long)(src_bytes / 16)); });
        snprintf(name, sizeof(name), "kernel pull 192/300, 4-byte loads, %d CTAs", ctas);
        time(name, dst_bytes, [&] { pull_rows_f32<<<ctas, 256>>>(h, d_take, rows, row_f, take_f); });
        snprintf(name, sizeof(name), "kernel pull 192/300, 16-byte loads, %d CTAs", ctas);
        time(name, dst_bytes, [&] { pull_rows_f32x4<<<ctas, 256>>>(h, d_take, rows, row_f, take_f); });
    }
    // spot check of the last variant
    float *chk = (float *)malloc(dst_bytes);
    CK(cudaMemcpy(chk, d_take, dst_bytes, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int r = 0; r < rows; r += 997) for (int c = 0; c < take_f; c++) bad += chk[(size_t)r * take_f + c] != h[(size_t)r * row_f + c];
    printf("check: %lld mismatches\n", bad);
    return bad != 0;
}
