// ugf_sort.cuh — parcel -> cell occupancy: counting sort with warp-aggregated atomics and a warp-cooperative
// per-cell index sort that makes the result stable (= independent of atomic order).
//
// Replaces CloudWithModels::buildCellOccupancy (CWM/CloudWithModels/CloudWithModels.C:110-138), whose in-cell
// order is cloud-list order.  Contract (SURVEY §7 hard part 7): CSR offsets + parcel ids, ids ascending
// within a cell w.r.t. the current device array order; deleted / migrating parcels (cell < 0) are dropped.
//
// Pipeline: histogram (fused into move_kernel) -> exclusive scan (3 small kernels over nCells ints)
//           -> index scatter (4 B read + 4 B write per parcel) -> per-cell segment sort (in place, skipped
//           when the segment is already ascending, which is the common case because the array is cell-major
//           from the previous step).  The payload is moved once, by the cell kernel's gather.
#pragma once
#include "ugf_common.cuh"

namespace ugf {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /* >= 33 ints */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        int w = lane < nw ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    const int res = smem[wid] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int* __restrict__ in, int n, int* __restrict__ blockSums) {
    __shared__ int sm[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) s += in[base + k];
    int total;
    block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

// single block: exclusive scan of blockSums in place; writes the grand total to *total and, if given, to
// *dN (the live parcel count becomes the array length after the gather).
__global__ void __launch_bounds__(SCAN_THREADS) scan_top_kernel(int* blockSums, int nb, int* total, long long* dN) {
    __shared__ int sm[33];
    int carry = 0;
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        const int idx = base + threadIdx.x;
        const int v = idx < nb ? blockSums[idx] : 0;
        int t;
        const int ex = block_exclusive_scan(v, &t, sm);
        if (idx < nb) blockSums[idx] = carry + ex;
        carry += t;
    }
    if (threadIdx.x == 0) {
        *total = carry;
        if (dN) *dN = carry;
    }
}

// offsets[i] = exclusive prefix of counts; counts are zeroed (they become the scatter cursors);
// offsets[n] = total.
__global__ void __launch_bounds__(SCAN_THREADS) scan_final_kernel(int* __restrict__ counts, int n, const int* __restrict__ blockSums,
                                                                 const int* __restrict__ total, int* __restrict__ offsets) {
    __shared__ int sm[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? counts[base + k] : 0; s += v[k]; }
    int t;
    int ex = block_exclusive_scan(s, &t, sm) + blockSums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) { offsets[base + k] = ex; counts[base + k] = 0; }
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n] = *total;
}

// perm[offsets[cell] + slot] = i, slot claimed through the per-cell cursor (lanes of a warp that share a
// cell claim a contiguous run with one atomic and keep their relative order).
__global__ void __launch_bounds__(256) scatter_index_kernel(const int* __restrict__ cell, const long long* __restrict__ dN,
                                                            const int* __restrict__ offsets, int* __restrict__ cursor,
                                                            int* __restrict__ perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = *dN;
    int c = -1;
    if (i < n) c = cell[i];
    const bool live = c >= 0;
    const unsigned liveMask = __ballot_sync(0xffffffffu, live);
    if (live) {
        const unsigned peers = __match_any_sync(liveMask, c);
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cursor[c], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        perm[offsets[c] + base + rank] = (int)i;
    }
}

// ---- per-cell segment sort ---------------------------------------------------------------------------
// Ascending-only bitonic network (partner = i ^ (k-1) for the first sub-step of a merge, then i ^ j): the
// lower index always takes the minimum, so virtual +inf padding beyond n never moves.
constexpr int SEG_THREADS = 256;
constexpr int SEG_SMEM_INTS = 1024;  // per-warp staging for segments of 33..1024 ids

__device__ __forceinline__ int warp_sort32(int v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        {
            const int mask = k - 1;
            const int o = __shfl_xor_sync(0xffffffffu, v, mask);
            v = ((lane ^ mask) > lane) ? min(v, o) : max(v, o);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(0xffffffffu, v, j);
            v = ((lane ^ j) > lane) ? min(v, o) : max(v, o);
        }
    }
    return v;
}

// generic in-memory version for one warp; a may be shared or global memory, indices >= n are +inf
__device__ inline void warp_sort_mem(int* a, int n, int lane) {
    int m = 1;
    while (m < n) m <<= 1;
    auto substep = [&](int mask) {
        for (int i = lane; i < m; i += 32) {
            const int p = i ^ mask;
            if (p > i && p < n) {
                const int x = a[i], y = a[p];
                if (x > y) { a[i] = y; a[p] = x; }
            }
        }
        __syncwarp();
    };
    for (int k = 2; k <= m; k <<= 1) {
        substep(k - 1);
        for (int j = k >> 2; j > 0; j >>= 1) substep(j);
    }
}

__global__ void __launch_bounds__(SEG_THREADS) segment_sort_kernel(const int* __restrict__ offsets, int nCells, int* __restrict__ perm) {
    __shared__ int stage[(SEG_THREADS / 32) * SEG_SMEM_INTS];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warpsTotal = gridDim.x * (SEG_THREADS / 32);
    int* sm = stage + wib * SEG_SMEM_INTS;
    for (int c = blockIdx.x * (SEG_THREADS / 32) + wib; c < nCells; c += warpsTotal) {
        const int beg = offsets[c];
        const int n = offsets[c + 1] - beg;
        if (n <= 1) continue;
        int* seg = perm + beg;
        if (n <= 32) {
            int v = lane < n ? seg[lane] : 0x7fffffff;
            const int prev = __shfl_up_sync(0xffffffffu, v, 1);
            const bool bad = lane > 0 && lane < n && prev > v;
            if (!__any_sync(0xffffffffu, bad)) continue;
            v = warp_sort32(v, lane);
            if (lane < n) seg[lane] = v;
            continue;
        }
        bool bad = false;
        for (int i = lane + 1; i < n; i += 32) bad |= seg[i - 1] > seg[i];
        if (!__any_sync(0xffffffffu, bad)) continue;
        if (n <= SEG_SMEM_INTS) {
            for (int i = lane; i < n; i += 32) sm[i] = seg[i];
            __syncwarp();
            warp_sort_mem(sm, n, lane);
            for (int i = lane; i < n; i += 32) seg[i] = sm[i];
            __syncwarp();
        } else {
            __syncwarp();
            warp_sort_mem(seg, n, lane);  // giant cells: in global memory (correct, slow)
        }
    }
}

// Standalone payload gather (ugf_reorder): out[j] = in[perm[j]].  The per-step path fuses this into cell_kernel.
template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(256) reorder_kernel(ParcelBuf in, ParcelBuf out, const int* __restrict__ perm, const int* __restrict__ total) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *total) return;
    const int s = perm[j];
    out.x[j] = in.x[s]; out.y[j] = in.y[s]; out.z[j] = in.z[s];
    out.ux[j] = in.ux[s]; out.uy[j] = in.uy[s]; out.uz[j] = in.uz[s];
    out.cell[j] = in.cell[s];
    if (HAS_ROT) out.erot[j] = in.erot[s];
    if (MULTI) out.type[j] = in.type[s];
}

}  // namespace ugf
