// gs_binning.cu -- stage 2 of the forward: per-tile lists of Gaussians in (depth, index) order.
//
// Semantics: SURVEY.md Appendix A "Binning" (upstream K2-K5).  All views of the call are binned by ONE
// scan / sort: the key is ((view * tiles + tile) << 32) | fp32 bits of camera-space depth, the value the
// Gaussian's index in its scene; a stable LSD radix sort then orders every tile by (depth, index) exactly as
// upstream's emission-order + stable sort does.  Only key bits [0, 32 + log2(V * tiles)) are sorted.
// The scan and the radix sort are CUB library calls (library code, like cuBLAS for a plain GEMM); the
// duplicate/range kernels are ours.  DESIGN.md section 5.2 has the traffic accounting.
#include <cub/cub.cuh>

#include "gs_common.cuh"

namespace {

__global__ void k_duplicate(const DevCfg c, const float4 *__restrict__ rec2, const uint32_t *__restrict__ tiles_touched,
                            const ushort4 *__restrict__ rects, const uint32_t *__restrict__ offsets,
                            uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)c.V * c.P;
    if (idx >= n) return;
    if (tiles_touched[idx] == 0) return;
    uint32_t off = idx == 0 ? 0u : offsets[idx - 1];
    const int v = (int)(idx / c.P);
    const uint32_t i = (uint32_t)(idx - (size_t)v * c.P);
    const ushort4 r = rects[idx];
    const uint32_t dbits = __float_as_uint(rec2[idx].y);
    const uint32_t tbase = (uint32_t)v * (uint32_t)c.ntiles;
    for (int y = r.y; y < r.w; y++)
        for (int x = r.x; x < r.z; x++) {
            keys[off] = ((uint64_t)(tbase + (uint32_t)(y * c.gx + x)) << 32) | dbits;
            vals[off] = i;
            off++;
        }
}

__global__ void k_ranges(int64_t D, const uint64_t *__restrict__ keys, uint2 *__restrict__ ranges) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D) return;
    const uint32_t t = (uint32_t)(keys[idx] >> 32);
    if (idx == 0 || t != (uint32_t)(keys[idx - 1] >> 32)) ranges[t].x = (uint32_t)idx;
    if (idx == D - 1 || t != (uint32_t)(keys[idx + 1] >> 32)) ranges[t].y = (uint32_t)(idx + 1);
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int sort_bits(const DevCfg &c) {
    int tb = 0;
    while ((1ll << tb) < (long long)c.V * c.ntiles) tb++;
    return 32 + tb;
}

}  // namespace

size_t bin_scan_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    return align256(bytes);
}

int bin_scan(const DevCfg &c, const uint32_t *tiles_touched, uint32_t *offsets, void *temp, size_t temp_bytes,
             cudaStream_t st) {
    const int64_t n = (int64_t)c.V * c.P;
    if (n == 0) return GS_OK;
    GS_CUDA_OK(cub::DeviceScan::InclusiveSum(temp, temp_bytes, tiles_touched, offsets, (int)n, st));
    return GS_OK;
}

// scratch = keys_in[D] | keys_out[D] | vals_in[D] | cub temp
size_t bin_scratch_bytes(const DevCfg &c, int64_t D) {
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)D, 0, sort_bits(c));
    return 2 * align256((size_t)D * 8) + align256((size_t)D * 4) + align256(temp);
}

int bin_sort(const DevCfg &c, int64_t D, const float4 *rec2, const uint32_t *tiles_touched, const ushort4 *rects,
             const uint32_t *offsets, void *scratch, size_t scratch_bytes, uint32_t *point_list, uint2 *ranges,
             cudaStream_t st) {
    GS_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)c.V * c.ntiles, st));
    if (D == 0) return GS_OK;
    unsigned char *p = static_cast<unsigned char *>(scratch);
    uint64_t *keys_in = reinterpret_cast<uint64_t *>(p);
    p += align256((size_t)D * 8);
    uint64_t *keys_out = reinterpret_cast<uint64_t *>(p);
    p += align256((size_t)D * 8);
    uint32_t *vals_in = reinterpret_cast<uint32_t *>(p);
    p += align256((size_t)D * 4);
    void *temp = p;
    size_t temp_bytes = scratch_bytes - (size_t)(p - static_cast<unsigned char *>(scratch));

    const size_t n = (size_t)c.V * c.P;
    k_duplicate<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c, rec2, tiles_touched, rects, offsets, keys_in, vals_in);
    GS_CUDA_OK(cudaGetLastError());
    GS_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, point_list, (int)D, 0,
                                               sort_bits(c), st));
    k_ranges<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(D, keys_out, ranges);
    GS_CUDA_OK(cudaGetLastError());
    return GS_OK;
}
