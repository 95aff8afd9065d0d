// gs_tile_sort.cuh -- shared-memory sort of one (view, tile) bucket, used by k_tile_sort (gs_binning.cu) and by the
// fused sort+composite forward kernel (gs_composite_fwd.cu).
#pragma once
#include <cub/block/block_merge_sort.cuh>

#include "gs_common.cuh"

// Bytes of shared memory sort_bucket_merge<THREADS, ITEMS> needs.
template <int THREADS, int ITEMS>
constexpr size_t tile_sort_smem_bytes() {
    return sizeof(typename cub::BlockMergeSort<uint64_t, THREADS, ITEMS>::TempStorage);
}

// 64-bit merge sort of one bucket: (depth_bits << 32 | index) is a total order, so the arbitrary arrival order of
// the bucket does not matter.  All THREADS threads of the CTA call it; dst receives the indices in order.
// `load(i)` returns key i of the bucket (i < n): a plain array for the exact-capacity path, a gather over the tile's
// sub-buckets for the speculative-capacity path.
template <int THREADS, int ITEMS, typename Load>
__device__ __forceinline__ void sort_bucket_merge(Load load, uint32_t *__restrict__ dst, uint32_t n, void *smem) {
    using Sort = cub::BlockMergeSort<uint64_t, THREADS, ITEMS>;
    typename Sort::TempStorage &tmp = *reinterpret_cast<typename Sort::TempStorage *>(smem);
    uint64_t keys[ITEMS];
    const uint32_t base = threadIdx.x * ITEMS;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) keys[k] = base + k < n ? load(base + k) : ~0ull;
    Sort(tmp).Sort(keys, [](const uint64_t &a, const uint64_t &b) { return a < b; });
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
        if (base + k < n) dst[base + k] = (uint32_t)keys[k];
}

// Sorts a bucket of n <= THREADS * MAX_ITEMS keys with the cheapest instantiation (ITEMS in MAX/8, MAX/2, MAX) that
// fits it.
template <int THREADS, int MAX_ITEMS, typename Load>
__device__ __forceinline__ void sort_bucket_dispatch(Load load, uint32_t *__restrict__ dst, uint32_t n, void *smem) {
    constexpr int LO = MAX_ITEMS >= 8 ? MAX_ITEMS / 8 : 1, MID = MAX_ITEMS >= 2 ? MAX_ITEMS / 2 : 1;
    if (n <= (uint32_t)THREADS * LO) sort_bucket_merge<THREADS, LO>(load, dst, n, smem);
    else if (n <= (uint32_t)THREADS * MID) sort_bucket_merge<THREADS, MID>(load, dst, n, smem);
    else sort_bucket_merge<THREADS, MAX_ITEMS>(load, dst, n, smem);
}
