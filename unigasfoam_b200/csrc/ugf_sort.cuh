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
//           from the previous step; in the fused step this happens inside the cell kernel, on the ids it has
//           just loaded).  The payload is moved once, by the cell kernel's gather.
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
__global__ void __launch_bounds__(SCAN_THREADS) scan_top_kernel(int* blockSums, int nb, int* total, long long* dN, long long capacity, int* err) {
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
        if (carry > capacity) {  // clones (cell weighting) outgrew the parcel arrays: report, and leave an empty occupancy
            if (err) atomicCAS(err, 0, 1);  // behind so that no later kernel writes past the arrays
            carry = 0;
        }
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
    const bool dead = (*total == 0);  // empty cloud, or the capacity overflow flagged by scan_top_kernel
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) { offsets[base + k] = dead ? 0 : ex; counts[base + k] = 0; }
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n] = *total;
}

// The occupancy build's variant without the single-block pass in between: every block sums the raw block totals
// written by scan_reduce_kernel itself (a few thousand ints at most), so the scan is two launches instead of three.
// The last block publishes the grand total and raises the capacity flag (cell-weighting clones can outgrow the arrays:
// an empty occupancy is left behind so that no later kernel writes past them).
__global__ void __launch_bounds__(SCAN_THREADS) scan_final_self_kernel(int* __restrict__ counts, int n, const int* __restrict__ blockSums, int nb,
                                                                      int* __restrict__ total, int* __restrict__ offsets, long long capacity,
                                                                      int* err) {
    __shared__ int sm[33];
    int pre = 0, tot = 0;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const int b = blockSums[j];
        tot += b;
        if (j < blockIdx.x) pre += b;
    }
    int sPre, sTot;
    block_exclusive_scan(pre, &sPre, sm);
    block_exclusive_scan(tot, &sTot, sm);
    const int grand = sTot;
    const bool dead = grand > capacity;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? counts[base + k] : 0; s += v[k]; }
    int t;
    int ex = block_exclusive_scan(s, &t, sm) + sPre;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) { offsets[base + k] = dead ? 0 : ex; counts[base + k] = 0; }
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        offsets[n] = dead ? 0 : grand;
        *total = dead ? 0 : grand;
        if (dead && err) atomicCAS(err, 0, 1);
    }
}

// Lanes of a warp that hold the same key in a run of consecutive lanes (the common case: parcels are nearly
// cell-major) are aggregated without match.any: head = first lane of the run, cnt = its length, rank = position
// in it.  Equal keys in separate runs simply issue separate atomics.
__device__ __forceinline__ void warp_runs(int key, int lane, int& head, int& cnt, int& rank) {
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
    const int end = above ? __ffs(above) - 1 : 32;
    cnt = end - head;
    rank = lane - head;
}

// perm[offsets[cell] + slot] = i, slot claimed through the per-cell cursor (lanes of a warp that share a
// cell claim a contiguous run with one atomic and keep their relative order).  A warp takes SCAT_ROWS rows of 32
// consecutive parcels and keeps all their load -> atomic -> store chains in flight together (the kernel is
// bound by the latency of that chain, not by its 8 B/parcel).
constexpr int SCAT_ROWS = 4;
__global__ void __launch_bounds__(256) scatter_index_kernel(const int* __restrict__ cell, const long long* __restrict__ dN,
                                                            const int* __restrict__ offsets, int* __restrict__ cursor,
                                                            int* __restrict__ perm, const uint8_t* __restrict__ nclone,
                                                            long long capacity) {
    const int lane = threadIdx.x & 31;
    const long long wbase = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 * SCAT_ROWS) + lane;
    const long long n = *dN;
    int c[SCAT_ROWS], base[SCAT_ROWS], off[SCAT_ROWS], head[SCAT_ROWS], rank[SCAT_ROWS];
#pragma unroll
    for (int r = 0; r < SCAT_ROWS; ++r) {
        const long long i = wbase + r * 32;
        c[r] = (i < n) ? cell[i] : -1;
    }
#pragma unroll
    for (int r = 0; r < SCAT_ROWS; ++r) {
        int cnt;
        warp_runs(c[r], lane, head[r], cnt, rank[r]);
        base[r] = 0;
        off[r] = 0;
        if (c[r] >= 0) {
            if (rank[r] == 0) base[r] = atomicAdd(&cursor[c[r]], cnt);
            off[r] = offsets[c[r]];
        }
    }
#pragma unroll
    for (int r = 0; r < SCAT_ROWS; ++r) {
        const int b = __shfl_sync(0xffffffffu, base[r], head[r]);
        if (c[r] >= 0 && (long long)off[r] + b + rank[r] < capacity) perm[off[r] + b + rank[r]] = (int)(wbase + r * 32);
    }
    if (nclone) {  // cell weighting: a parcel with k clones claims k more slots of its cell for entries CLONE_FLAG | id
#pragma unroll
        for (int r = 0; r < SCAT_ROWS; ++r) {
            const long long i = wbase + r * 32;
            const int k = (c[r] >= 0) ? nclone[i] : 0;
            if (k) {
                const int b = atomicAdd(&cursor[c[r]], k);
                for (int j = 0; j < k; ++j)
                    if ((long long)off[r] + b + j < capacity) perm[off[r] + b + j] = CLONE_FLAG | (int)i;
            }
        }
    }
}

// ---- per-cell segment sort ---------------------------------------------------------------------------
// Ascending-only bitonic network (partner = i ^ (k-1) for the first sub-step of a merge, then i ^ j): the
// lower index always takes the minimum, so virtual +inf padding beyond n never moves.
constexpr int SEG_THREADS = 256;
constexpr int SEG_SMEM_INTS = 512;  // per-warp staging: a chunk of cells, or one segment of 33..512 ids

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

// One warp per chunk of SEG_CHUNK consecutive cells, SEG_LPC lanes per cell: every cell's id segment is checked
// for ascending order by its own lanes, all cells of the chunk in parallel; the cells found out of order are then
// sorted one after another by the whole warp (bitonic network in registers for <= 32 ids, in shared memory up to
// SEG_SMEM_INTS, in global memory beyond).  The scatter claims slots in atomic order and a cell's parcels come
// from several warps (its own previous range plus arrivals from neighbour cells whose ranges lie far away in the
// array), so a large share of the segments does need the sort.  Measured alternatives that were slower on the
// bench workload: rank sort by the cell's lanes (98 us vs 82 us), ranking only the runs of consecutive ids
// (160 us: runs are ~2 ids long), ordering inside the streaming cell kernel (+83 us there).
constexpr int SEG_CHUNK = 8;
constexpr int SEG_LPC = 32 / SEG_CHUNK;

__device__ inline void warp_sort_segment(int* seg, int n, int* sm, int lane) {
    if (n <= 32) {
        int v = lane < n ? seg[lane] : 0x7fffffff;
        v = warp_sort32(v, lane);
        if (lane < n) seg[lane] = v;
    } else if (n <= SEG_SMEM_INTS) {
        for (int i = lane; i < n; i += 32) sm[i] = seg[i];
        __syncwarp();
        warp_sort_mem(sm, n, lane);
        for (int i = lane; i < n; i += 32) seg[i] = sm[i];
    } else {
        warp_sort_mem(seg, n, lane);  // giant cells: in global memory (correct, slow)
    }
    __syncwarp();
}

__global__ void __launch_bounds__(SEG_THREADS) segment_sort_kernel(const int* __restrict__ offsets, int nCells, int* __restrict__ perm,
                                                                  int* __restrict__ cursor) {
    __shared__ int stage[(SEG_THREADS / 32) * SEG_SMEM_INTS];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    int* sm = stage + wib * SEG_SMEM_INTS;
    const int chunk = blockIdx.x * (SEG_THREADS / 32) + wib;
    const int c0 = chunk * SEG_CHUNK;
    if (c0 >= nCells) return;
    const int nc = min(SEG_CHUNK, nCells - c0);
    const int offv = (lane <= nc) ? offsets[c0 + lane] : 0;
    if (lane < nc) cursor[c0 + lane] = 0;  // the scatter is done with the cursors: leave the histogram array zeroed for the next move
    const int g = lane / SEG_LPC, q = lane % SEG_LPC;
    const int cb = __shfl_sync(0xffffffffu, offv, g);
    const int ce = __shfl_sync(0xffffffffu, offv, (g + 1) & 31);
    const int n = (g < nc) ? ce - cb : 0;
    bool bad = false;
    // lane q checks the pairs (i-1, i), i = q+1, q+1+SEG_LPC, ...
    int prev = (q < n) ? perm[cb + q] : 0;
    for (int i = q + 1; i < n; i += SEG_LPC) {
        const int v = perm[cb + i];
        const int nxt = (i + SEG_LPC - 1 < n) ? perm[cb + i + SEG_LPC - 1] : 0;
        bad |= prev > v;
        prev = nxt;
    }
    unsigned badMask = __ballot_sync(0xffffffffu, bad);
    while (badMask) {
        const int gi = (__ffs(badMask) - 1) / SEG_LPC;
        badMask &= ~(((1u << SEG_LPC) - 1u) << (gi * SEG_LPC));
        const int b = __shfl_sync(0xffffffffu, offv, gi);
        const int e = __shfl_sync(0xffffffffu, offv, gi + 1);
        warp_sort_segment(perm + b, e - b, sm, lane);
    }
}

// Standalone payload gather (ugf_reorder): out[j] = in[perm[j]].  The per-step path fuses this into cell_kernel.
template <bool HAS_ROT, bool MULTI>
__global__ void __launch_bounds__(256) reorder_kernel(ParcelBuf in, ParcelBuf out, const int* __restrict__ perm, const int* __restrict__ total) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *total) return;
    const int s = perm[j] & CLONE_MASK;
    out.x[j] = in.x[s]; out.y[j] = in.y[s]; out.z[j] = in.z[s];
    out.ux[j] = in.ux[s]; out.uy[j] = in.uy[s]; out.uz[j] = in.uz[s];
    out.cell[j] = in.cell[s];
    if (HAS_ROT) out.erot[j] = in.erot[s];
    if (MULTI) out.type[j] = in.type[s];
    if (in.vib) out.vib[j] = in.vib[s];
    if (in.elev) out.elev[j] = in.elev[s];
}

}  // namespace ugf
