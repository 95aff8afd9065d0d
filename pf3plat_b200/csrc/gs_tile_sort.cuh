// gs_tile_sort.cuh -- shared-memory sorts of one (view, tile) bucket for k_tile_sort / k_tile_sort_spec
// (gs_binning.cu): a 64-bit merge sort (cub::BlockMergeSort).
#pragma once
#include <cub/block/block_merge_sort.cuh>

#include "gs_common.cuh"

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

// Tried and dropped: a hand-written shared-memory LSD radix sort on the depth bits minus the tile's minimum (8-bit
// digits, per-warp match_any ranking, equal-depth runs fixed up by insertion on the index).  C2 tile sort: 0.367 ms
// as 1024 threads x 4 keys (53 registers -> one CTA per SM, ~20 CTA-wide barriers per tile fully exposed), 0.26 ms as
// 512 x 8 (two CTAs per SM) -- against 0.253 ms for the merge sort below, which needs no tie handling.

// Bytes of shared memory sort_bucket_dispatch<THREADS, MAX_ITEMS> needs.
template <int THREADS, int MAX_ITEMS>
constexpr size_t tile_sort_smem_bytes() {
    const size_t m = sizeof(typename cub::BlockMergeSort<uint64_t, THREADS, MAX_ITEMS>::TempStorage);
    return m;
}

// Sorts a bucket of n <= THREADS * MAX_ITEMS keys with the cheapest instantiation (ITEMS in MAX/8, MAX/2, MAX) that fits.
template <int THREADS, int MAX_ITEMS, typename Load>
__device__ __forceinline__ void sort_bucket_dispatch(Load load, uint32_t *__restrict__ dst, uint32_t n, void *smem) {
    constexpr int LO = MAX_ITEMS >= 8 ? MAX_ITEMS / 8 : 1, MID = MAX_ITEMS >= 2 ? MAX_ITEMS / 2 : 1;
    if (n <= (uint32_t)THREADS * LO) sort_bucket_merge<THREADS, LO>(load, dst, n, smem);
    else if (n <= (uint32_t)THREADS * MID) sort_bucket_merge<THREADS, MID>(load, dst, n, smem);
    else sort_bucket_merge<THREADS, MAX_ITEMS>(load, dst, n, smem);
}
