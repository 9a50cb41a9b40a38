// grid_uf.cuh -- K1/K3/K4/K5: voxel-hash neighbour grid and lock-free union-find connected components.
//
// Replaces pcl::search::KdTree + pcl::EuclideanClusterExtraction::extract (reference call site
// MOT.cpp:472-488).  The result is the partition into connected components of the graph
//     { (i,j) : fp32 ((dx*dx)+dy*dy)+dz*dz  <  r2 },  r2 = (float)((double)tol * (double)tol)
// which is exactly what PCL's BFS over FLANN radius searches computes.
//
// Grid.  Coarse cell edge h = tol*(1+2^-10) (north_star's "cell size equal to the tolerance" plus a safety
// epsilon so fp32 rounding of the predicate can never put a qualifying pair two coarse cells apart); each
// coarse cell is split 2x2x2 into fine cells of edge e = h/2.  A fine cell's diagonal is 0.867*tol, so all of
// its points are mutually adjacent: a fine cell is a clique and becomes ONE union-find node -- no pair tests
// inside it.  Between two fine cells a single witness pair is enough to merge them, and a pair of cells that
// already share a root is skipped without looking at any point.  This is what keeps dense frames (hundreds
// of points per cell on a LiDAR-sampled surface) near linear instead of quadratic in cell occupancy.
//
// Key layout (low to high bits): 3 fine bits (fz fy fx) | cx | cy | cz | frame id (batch mode).
// Points are radix sorted by key, so the fine cells of one coarse cell are contiguous; an open-addressing hash maps a
// coarse key to its coarse-cell index, and k_coarse_records resolves every coarse cell's forward half stencil (13
// neighbours) into one 16-int row.  Union-find work is dealt per coarse cell: k_uf_sparse stages the cell's forward
// neighbourhood in shared memory (TMA) and sweeps it by brute force with the exact predicate; neighbourhoods too large
// for the tile go to k_uf_dense, which looks for one witness pair per (fine cell, fine cell) candidate.  The first
// generation (k_uf_pairs: a warp per fine cell probing 27 coarse neighbours) is kept selectable for A/B (MOT_UF_MODE=0).
#pragma once
#include "common.cuh"

namespace mot {

// d_counts layout (ints)
enum { CNT_FINE = 0, CNT_COARSE = 1, CNT_K = 2, CNT_TOTAL = 3, CNT_FLAGS = 4, CNT_M = 5, CNT_DENSE = 6, CNT_HB = 8, CNT_N = 24 };

struct GridCodec {
    double minx, miny, minz, inv_e;
    int bx, by, bz;          // coarse-coordinate bits per axis
    int ncx, ncy, ncz;       // coarse cells per axis
    int nfx, nfy, nfz;       // fine cells per axis
    int n_frames;            // 1 unless batch mode
};

template <typename KT>
__device__ __forceinline__ KT key_compose(const GridCodec& g, int frame, int ix, int iy, int iz) {
    KT ck = (KT)frame;
    ck = (ck << g.bz) | (KT)(iz >> 1);
    ck = (ck << g.by) | (KT)(iy >> 1);
    ck = (ck << g.bx) | (KT)(ix >> 1);
    return (ck << 3) | (KT)(((iz & 1) << 2) | ((iy & 1) << 1) | (ix & 1));
}
template <typename KT>
__device__ __forceinline__ void key_decode(const GridCodec& g, KT key, int& frame, int& ix, int& iy, int& iz) {
    const int f = (int)(key & 7);
    KT ck = key >> 3;
    const int cx = (int)(ck & (((KT)1 << g.bx) - 1)); ck >>= g.bx;
    const int cy = (int)(ck & (((KT)1 << g.by) - 1)); ck >>= g.by;
    const int cz = (int)(ck & (((KT)1 << g.bz) - 1)); ck >>= g.bz;
    frame = (int)ck;
    ix = 2 * cx + (f & 1);
    iy = 2 * cy + ((f >> 1) & 1);
    iz = 2 * cz + (f >> 2);
}
template <typename KT>
__device__ __forceinline__ KT coarse_compose(const GridCodec& g, int frame, int cx, int cy, int cz) {
    KT ck = (KT)frame;
    ck = (ck << g.bz) | (KT)cz;
    ck = (ck << g.by) | (KT)cy;
    ck = (ck << g.bx) | (KT)cx;
    return ck;
}

// frame of point i in batch mode: largest f with frame_offsets[f] <= i
__device__ __forceinline__ int frame_of(const int* __restrict__ frame_offsets, int n_frames, int i) {
    int lo = 0, hi = n_frames - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (frame_offsets[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// K1: voxel key per point (fp64 cell coordinates: an fp32 product could misplace a point by a cell).
// A point outside the grid (flag 32) or with a non-finite coordinate (flag 64) is reported: with a grid planned from this
// call's own bounding box neither can happen; with the handle's SPECULATIVE plan (the grid of its previous call, no bounding
// box pass and no host round trip before the keys) the first one tells the host to plan again and rerun.
__device__ __forceinline__ void cell_coords_checked(const GridCodec& g, const float4& p, int& ix, int& iy, int& iz, int* __restrict__ flags) {
    ix = __double2int_rd(__dmul_rn(__dsub_rn((double)p.x, g.minx), g.inv_e));
    iy = __double2int_rd(__dmul_rn(__dsub_rn((double)p.y, g.miny), g.inv_e));
    iz = __double2int_rd(__dmul_rn(__dsub_rn((double)p.z, g.minz), g.inv_e));
    const bool finite = (fabsf(p.x) < INFINITY) && (fabsf(p.y) < INFINITY) && (fabsf(p.z) < INFINITY);
    if (!finite) atomicOr(flags, 64);
    else if (ix < 0 || iy < 0 || iz < 0 || ix >= g.nfx || iy >= g.nfy || iz >= g.nfz) atomicOr(flags, 32);
    ix = min(max(ix, 0), g.nfx - 1);
    iy = min(max(iy, 0), g.nfy - 1);
    iz = min(max(iz, 0), g.nfz - 1);
}

template <typename KT>
__global__ void __launch_bounds__(256) k_cell_keys(const float4* __restrict__ pts, int m, GridCodec g,
                                                    const int* __restrict__ frame_offsets, KT* __restrict__ keys, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float4 p = ld_stream(pts + i);
    int ix, iy, iz;
    cell_coords_checked(g, p, ix, iy, iz, flags);
    const int frame = g.n_frames > 1 ? frame_of(frame_offsets, g.n_frames, i) : 0;
    keys[i] = key_compose<KT>(g, frame, ix, iy, iz);
}

// ---- hash of coarse cells ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash_u(uint32_t k) { return k * 0x9E3779B1u; }
__device__ __forceinline__ unsigned hash_u(uint64_t k) { return (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 32); }

__device__ __forceinline__ uint32_t cas_key(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ uint64_t cas_key(uint64_t* p, uint64_t cmp, uint64_t v) {
    return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)v);
}

template <typename KT>
__device__ __forceinline__ void hash_insert(KT* hkeys, int* hvals, unsigned hmask, int hshift, KT ck, int val) {
    unsigned slot = hash_u(ck) >> hshift;
    for (unsigned probe = 0; probe <= hmask; ++probe) {
        const KT prev = cas_key(&hkeys[slot], ~(KT)0, ck);
        if (prev == ~(KT)0 || prev == ck) {
            hvals[slot] = val;
            return;
        }
        slot = (slot + 1) & hmask;
    }
}
template <typename KT>
__device__ __forceinline__ int hash_find(const KT* __restrict__ hkeys, const int* __restrict__ hvals, unsigned hmask, int hshift, KT ck) {
    unsigned slot = hash_u(ck) >> hshift;
    for (unsigned probe = 0; probe <= hmask; ++probe) {
        const KT k = hkeys[slot];
        if (k == ck) return hvals[slot];
        if (k == ~(KT)0) return -1;
        slot = (slot + 1) & hmask;
    }
    return -1;
}

// ---- K3: reorder + cell tables ---------------------------------------------------------------------------
constexpr int CELL_THREADS = 256;
constexpr int CELL_MAX_GRID = 592;
// k_cells_count / k_cells_write run with many more blocks than fit on the GPU at once: blocks start in index order, so the
// resident ones work on one neighbourhood of the sorted array and the random 16-byte point gathers stay inside an L2-sized
// window of the input (a batch of frames is far larger than L2; 592 chunks covering everything at once measured 1.55x slower)
constexpr int CELLW_MAX_GRID = 8192;

// counts[0][b] = fine-cell heads in block b's chunk, counts[1][b] = coarse-cell heads.
template <typename KT>
__global__ void __launch_bounds__(CELL_THREADS) k_cells_count(const KT* __restrict__ skeys, int m, int chunk, int* __restrict__ counts) {
    __shared__ int red[2][CELL_THREADS / 32];
    const int begin = blockIdx.x * chunk, end = min(m, begin + chunk);
    int nf = 0, nc = 0;
    for (int j = begin + threadIdx.x; j < end; j += CELL_THREADS) {
        const KT k = skeys[j];
        if (j == 0) { ++nf; ++nc; }
        else {
            const KT kp = skeys[j - 1];
            nf += (k != kp);
            nc += ((k >> 3) != (kp >> 3));
        }
    }
    nf = warp_sum(nf); nc = warp_sum(nc);
    if (lane_id() == 0) { red[0][warp_id()] = nf; red[1][warp_id()] = nc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, b = 0;
        for (int w = 0; w < CELL_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
        counts[blockIdx.x] = a;
        counts[gridDim.x + blockIdx.x] = b;
    }
}

// The coarse-cell hash is sized on the device from the head count k_cells_count just produced (2-4 slots per occupied
// coarse cell: short probe chains, and the table of a batch of LiDAR frames stays L2 resident -- 4-8 slots measured
// slower) and cleared here; every later kernel reads log2(size) from d_counts[CNT_HB].
template <typename KT>
__global__ void __launch_bounds__(256) k_hash_clear(const int* __restrict__ counts, int n_blocks, int hb_max, KT* __restrict__ hkeys,
                                                     int* __restrict__ d_counts, int* __restrict__ prefix) {
    __shared__ int scratch[36];
    if (blockIdx.x == 0) {
        // exclusive prefixes of the per-block head counts (fine cells, then coarse cells) for k_cells_write
        const int per = (n_blocks + 255) / 256;
        const int i0 = threadIdx.x * per, i1 = min(n_blocks, i0 + per);
        for (int arr = 0; arr < 2; ++arr) {
            const int* c = counts + arr * n_blocks;
            int* o = prefix + arr * n_blocks;
            int sum = 0;
            for (int i = i0; i < i1; ++i) sum += c[i];
            int total;
            int run = block_exclusive_scan(sum, scratch, &total);
            for (int i = i0; i < i1; ++i) {
                o[i] = run;
                run += c[i];
            }
        }
    }
    const int n_coarse = block_prefix_of(counts + n_blocks, n_blocks, scratch);
    int hb = 4;
    while (hb < hb_max && (1ll << hb) < 2ll * n_coarse) ++hb;
    if (blockIdx.x == 0 && threadIdx.x == 0) d_counts[CNT_HB] = hb;
    const size_t size = (size_t)1 << hb;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (size_t)gridDim.x * blockDim.x) hkeys[i] = ~(KT)0;
}

// d_counts layout (ints): [0] fine cells, [1] coarse cells, [2] kept clusters, [3] kept points, [4] flags


constexpr int CELLW_ITEMS = 4;  // consecutive sorted positions per thread: one block scan per 1024 points, four gathers in flight
constexpr int CELLW_TILE = CELL_THREADS * CELLW_ITEMS;

template <typename KT>
__global__ void __launch_bounds__(CELL_THREADS) k_cells_write(const KT* __restrict__ skeys, const uint32_t* __restrict__ svals,
                                                               const float4* __restrict__ pts, float4* __restrict__ spts, int m,
                                                               int chunk, const int* __restrict__ counts, int* __restrict__ fc_start,
                                                               int* __restrict__ cc_first,
                                                               int* __restrict__ parent, int* __restrict__ csize, int* __restrict__ cmin,
                                                               int* __restrict__ crank, int* __restrict__ d_counts,
                                                               KT* __restrict__ ckey_out /* sorted coarse keys */,
                                                               unsigned char* __restrict__ fcode_out /* child code of every fine cell, may be null */) {
    __shared__ int scratch[36];
    int fbase = counts[2 * gridDim.x + blockIdx.x];  // exclusive prefixes written by k_hash_clear behind the two count rows
    int cbase = counts[3 * gridDim.x + blockIdx.x];
    const int begin = blockIdx.x * chunk, end = min(m, begin + chunk);
    for (int tb = begin; tb < end; tb += CELLW_TILE) {
        const int j0 = tb + threadIdx.x * CELLW_ITEMS;
        KT k[CELLW_ITEMS];
        uint32_t sv[CELLW_ITEMS];
        float4 p[CELLW_ITEMS];
        unsigned fh = 0, ch = 0;  // bit i: position j0 + i starts a fine / coarse cell
        KT kp = 0;
        if (j0 > 0 && j0 < end) kp = skeys[j0 - 1];
#pragma unroll
        for (int i = 0; i < CELLW_ITEMS; ++i) {
            const int j = j0 + i;
            k[i] = 0;
            sv[i] = 0;
            if (j < end) {
                k[i] = skeys[j];
                sv[i] = svals[j];
            }
        }
#pragma unroll
        for (int i = 0; i < CELLW_ITEMS; ++i)
            if (j0 + i < end) p[i] = pts[sv[i]];  // the random 16-byte gathers of the tile, all in flight before the scan's barriers
#pragma unroll
        for (int i = 0; i < CELLW_ITEMS; ++i) {
            const int j = j0 + i;
            if (j < end) {
                const KT prev = i == 0 ? kp : k[i - 1];
                if (j == 0 || k[i] != prev) fh |= 1u << i;
                if (j == 0 || (k[i] >> 3) != (prev >> 3)) ch |= 1u << i;
            }
        }
        int total;
        const int excl = block_exclusive_scan((__popc(ch) << 16) | __popc(fh), scratch, &total);
        int f_run = fbase + (excl & 0xffff), c_run = cbase + (excl >> 16);
#pragma unroll
        for (int i = 0; i < CELLW_ITEMS; ++i) {
            const int j = j0 + i;
            if (j >= end) break;
            const bool is_fh = (fh >> i) & 1u, is_ch = (ch >> i) & 1u;
            f_run += is_fh;
            const int fi = f_run - 1;  // index of the fine cell containing j
            // sorted SoA point: (x, y, z, bits(fine cell id)) -- the union-find kernels get a point's cell for free
            p[i].w = __int_as_float(fi);
            spts[j] = p[i];
            if (is_fh) {
                fc_start[fi] = j;
                parent[fi] = fi;
                csize[fi] = 0;
                cmin[fi] = (int)sv[i];  // stable sort: the first point of a cell has its smallest original index
                crank[fi] = -1;
                if (fcode_out) fcode_out[fi] = (unsigned char)(k[i] & 7);
            }
            if (is_ch) {
                cc_first[c_run] = fi;
                ckey_out[c_run] = k[i] >> 3;
                ++c_run;
            }
        }
        fbase += total & 0xffff;
        cbase += total >> 16;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        fc_start[fbase] = m;
        cc_first[cbase] = fbase;
        d_counts[CNT_FINE] = fbase;
        d_counts[CNT_COARSE] = cbase;
    }
}

// coarse key -> coarse cell index.  One thread per coarse cell: in k_cells_write the insert was done by the one lane in ~10 that
// starts a coarse cell while the other lanes of the warp waited for its compare-and-swap (27 % of that kernel's stall samples).
template <typename KT>
__global__ void __launch_bounds__(256) k_hash_build(const KT* __restrict__ ckey, const int* __restrict__ d_counts, KT* __restrict__ hkeys,
                                                     int* __restrict__ hvals) {
    const int n_coarse = d_counts[CNT_COARSE];
    const int hb = d_counts[CNT_HB];
    const unsigned hmask = (1u << hb) - 1u;
    const int hshift = 32 - hb;
    for (int ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n_coarse; ci += gridDim.x * blockDim.x)
        hash_insert<KT>(hkeys, hvals, hmask, hshift, ckey[ci], ci);
}

// ---- union-find primitives (parents only ever decrease; atomicMin hooking + path halving) --------------
__device__ __forceinline__ int uf_find(int* parent, int x) {
    for (;;) {
        const int p = ld_cg(parent + x);
        if (p == x) return x;
        const int gp = ld_cg(parent + p);
        if (gp == p) return p;
        st_cg(parent + x, gp);  // path halving; a racing atomicMin is re-established by its own retry loop
        x = gp;
    }
}
__device__ __forceinline__ void uf_unite(int* parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }  // a = larger root, hooked under the smaller one
        const int old = atomicMin(parent + a, b);
        if (old == a) return;                          // a was still a root: done
        a = old;                                       // a had been hooked meanwhile: connect its old parent with b
    }
}

__device__ __forceinline__ bool coop_witness(const float4* __restrict__ spts, int a0, int a1, int b0, int b1, float r2) {
    const int lane = lane_id();
    {   // probe: 32 scattered (p, q) pairs in one step -- between densely sampled neighbours one of them almost always hits
        const int na = a1 - a0, nb = b1 - b0;
        const float4 p = __ldg(spts + a0 + (int)(((unsigned)lane * 2654435761u >> 8) % (unsigned)na));
        const float4 q = __ldg(spts + b0 + (int)(((unsigned)lane * 40503u + 17u) % (unsigned)nb));
        if (__any_sync(kFull, dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < r2)) return true;
    }
    for (int ia = a0; ia < a1; ia += 32) {
        const bool pv = ia + lane < a1;
        const float4 p = pv ? spts[ia + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int jb = b0; jb < b1; jb += 32) {
            const bool qv = jb + lane < b1;
            const float4 q = qv ? spts[jb + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            const int nq = min(32, b1 - jb);
            bool hit = false;
            for (int s = 0; s < nq; ++s) {
                const float qx = __shfl_sync(kFull, q.x, s), qy = __shfl_sync(kFull, q.y, s), qz = __shfl_sync(kFull, q.z, s);
                hit |= dist2_exact(p.x, p.y, p.z, qx, qy, qz) < r2;
            }
            if (__any_sync(kFull, hit && pv)) return true;
        }
    }
    return false;
}

constexpr int UF_THREADS = 256;
constexpr int UF_WARPS = UF_THREADS / 32;
constexpr int UF_MAX_CAND = 27 * 8;
constexpr int UF_SMALL_PAIR = 64;  // cell pairs with at most this many point pairs are searched by one lane

// PHASE 1 merges fine cells at Chebyshev distance <= 1, PHASE 2 the ring at distance 2.  Running the rings as
// two launches (with a compress in between) means nearly every ring-2 pair is already connected through
// ring 1 and is skipped by the root comparison without touching a point.
template <typename KT, int PHASE>
__global__ void __launch_bounds__(UF_THREADS) k_uf_pairs(const KT* __restrict__ skeys, const float4* __restrict__ spts,
                                                          const int* __restrict__ fc_start, const int* __restrict__ cc_first,
                                                          const KT* __restrict__ hkeys, const int* __restrict__ hvals, unsigned hmask,
                                                          int hshift, const int* __restrict__ d_counts, int* parent, GridCodec g, float r2) {
    __shared__ int cand[UF_WARPS][UF_MAX_CAND];
    {
        const int hb = d_counts[CNT_HB];
        hmask = (1u << hb) - 1u;
        hshift = 32 - hb;
    }
    const int n_fine = d_counts[CNT_FINE];
    const int lane = lane_id(), w = warp_id();
    const int n_warps = gridDim.x * UF_WARPS;
    for (int A = blockIdx.x * UF_WARPS + w; A < n_fine; A += n_warps) {
        const int a0 = fc_start[A], a1 = fc_start[A + 1];
        const KT key_a = skeys[a0];
        int frame, ixa, iya, iza;
        key_decode<KT>(g, key_a, frame, ixa, iya, iza);
        int nfirst = 0, ncnt = 0;
        if (lane < 27) {
            const int cx = (ixa >> 1) + (lane % 3) - 1, cy = (iya >> 1) + ((lane / 3) % 3) - 1, cz = (iza >> 1) + (lane / 9) - 1;
            if (cx >= 0 && cy >= 0 && cz >= 0 && cx < g.ncx && cy < g.ncy && cz < g.ncz) {
                const int ci = hash_find<KT>(hkeys, hvals, hmask, hshift, coarse_compose<KT>(g, frame, cx, cy, cz));
                if (ci >= 0) {
                    nfirst = cc_first[ci];
                    ncnt = cc_first[ci + 1] - nfirst;
                }
            }
        }
        const int incl = warp_inclusive_scan(ncnt);
        const int T = __shfl_sync(kFull, incl, 31);
        for (int k = 0; k < ncnt; ++k) cand[w][incl - ncnt + k] = nfirst + k;
        __syncwarp();
        for (int base = 0; base < T; base += 32) {
            const int t = base + lane;
            int B = -1, b0 = 0, b1 = 0;
            bool big = false;
            if (t < T) {
                B = cand[w][t];
                bool ok = B > A;  // every unordered pair of fine cells is handled once, by its smaller index
                if (ok) {
                    b0 = fc_start[B];
                    b1 = fc_start[B + 1];
                    int fb, ixb, iyb, izb;
                    key_decode<KT>(g, skeys[b0], fb, ixb, iyb, izb);
                    const int ring = max(max(abs(ixb - ixa), abs(iyb - iya)), abs(izb - iza));
                    ok = PHASE == 1 ? ring <= 1 : ring == 2;
                }
                if (ok) ok = uf_find(parent, A) != uf_find(parent, B);
                if (ok) {
                    const int na = a1 - a0, nb = b1 - b0;
                    if (na <= UF_SMALL_PAIR && nb <= UF_SMALL_PAIR && na * nb <= UF_SMALL_PAIR) {
                        bool hit = false;
                        for (int ia = a0; ia < a1 && !hit; ++ia) {
                            const float4 p = spts[ia];
                            for (int jb = b0; jb < b1; ++jb) {
                                const float4 q = spts[jb];
                                if (dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < r2) { hit = true; break; }
                            }
                        }
                        if (hit) uf_unite(parent, A, B);
                    } else {
                        big = true;
                    }
                }
            }
            unsigned bm = __ballot_sync(kFull, big);
            while (bm) {
                const int s = __ffs(bm) - 1;
                bm &= bm - 1;
                const int Bs = __shfl_sync(kFull, B, s), b0s = __shfl_sync(kFull, b0, s), b1s = __shfl_sync(kFull, b1, s);
                int same = 0;
                if (lane == 0) same = uf_find(parent, A) == uf_find(parent, Bs);
                same = __shfl_sync(kFull, same, 0);
                if (same) continue;
                const bool found = coop_witness(spts, a0, a1, b0s, b1s, r2);
                if (found && lane == 0) uf_unite(parent, A, Bs);
            }
        }
        __syncwarp();
    }
}

// ---- K4: one warp per COARSE cell (records + half stencil, then k_uf_sparse / k_uf_dense) ----------------------------
//
// Record of an occupied coarse cell: x = first sorted point, y = point count, z = first fine cell, w = bit mask
// of its occupied fine children (child code = fz<<2 | fy<<1 | fx).  The fine children are consecutive fine-cell
// ids in child-code order, so fine id = z + rank of the child inside the mask.
// Also resolves the half stencil once: nbr[ci*16 + c] = index of the coarse cell at stencil offset 13 + c
// (c = 0 is ci itself, -1 = unoccupied), so the union-find warps start from one coalesced row load instead of
// a key -> hash -> value -> record chain.  16 threads per coarse cell.
template <typename KT>
__global__ void __launch_bounds__(256) k_coarse_records(const KT* __restrict__ skeys, const int* __restrict__ fc_start,
                                                         const int* __restrict__ cc_first, const int* __restrict__ d_counts,
                                                         const KT* __restrict__ hkeys, const int* __restrict__ hvals, unsigned hmask, int hshift,
                                                         GridCodec g, int4* __restrict__ crec, int* __restrict__ nbr) {
    const int n_coarse = d_counts[CNT_COARSE];
    {
        const int hb = d_counts[CNT_HB];
        hmask = (1u << hb) - 1u;
        hshift = 32 - hb;
    }
    const int sub = threadIdx.x & 15;
    for (int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; ci < n_coarse; ci += (gridDim.x * blockDim.x) >> 4) {
        const int f0 = cc_first[ci];
        const int p0 = fc_start[f0];
        KT ck = skeys[p0] >> 3;
        if (sub == 0) {
            const int f1 = cc_first[ci + 1];
            const int p1 = fc_start[f1];
            unsigned mask = 0;
            for (int f = f0; f < f1; ++f) mask |= 1u << (unsigned)(skeys[fc_start[f]] & 7);
            crec[ci] = make_int4(p0, p1 - p0, f0, (int)mask);
            nbr[ci * 16] = ci;
        } else {
            int nci = -1;
            if (sub < 14) {
                const int cxa = (int)(ck & (((KT)1 << g.bx) - 1)); ck >>= g.bx;
                const int cya = (int)(ck & (((KT)1 << g.by) - 1)); ck >>= g.by;
                const int cza = (int)(ck & (((KT)1 << g.bz) - 1)); ck >>= g.bz;
                const int frame = (int)ck;
                const int sidx = 13 + sub;
                const int cx = cxa + (sidx % 3) - 1, cy = cya + ((sidx / 3) % 3) - 1, cz = cza + (sidx / 9) - 1;
                if (cx >= 0 && cy >= 0 && cz >= 0 && cx < g.ncx && cy < g.ncy && cz < g.ncz)
                    nci = hash_find<KT>(hkeys, hvals, hmask, hshift, coarse_compose<KT>(g, frame, cx, cy, cz));
            }
            nbr[ci * 16 + sub] = nci;
        }
    }
}

constexpr int UFC_THREADS = 256;
constexpr int UFC_WARPS = UFC_THREADS / 32;
constexpr int UFC_TILE_PTS = 256;   // points of the forward neighbourhood staged per warp (4 KB)
constexpr int UFC_CELLS = 14;       // half stencil: the coarse cell itself + its 13 "forward" neighbours
constexpr int UFC_NODES = UFC_CELLS * 8;
#ifndef UFC_MIN_BLOCKS
#define UFC_MIN_BLOCKS 5
#endif
constexpr int UFC_BRUTE_TESTS = 4096;  // brute-force sweep only while (own points) x (neighbourhood points) stays below this
struct __align__(16) UfcWarpSmem {
    float4 tile[UFC_TILE_PTS];
    int cstart[16];            // first sorted point of cell c
    int coff[16];              // position of cell c's first point in the tile; coff[14] = coff[15] = total
    int ffirst[16];            // first fine id of cell c
    unsigned char attach[UFC_NODES];  // for neighbour node x = c*8 + j: an own child adjacent to it (0xff = none)
    uint64_t bar;
};

// Sparse tasks.  One warp per occupied coarse cell A (half stencil: cell 0 = A, cells 1..13 = the neighbours that
// follow A in (z, y, x) order, so every unordered pair of adjacent coarse cells belongs to exactly one warp).
// The forward neighbourhood (<= UFC_TILE_PTS points, the normal case on LiDAR frames) is staged in shared memory
// with TMA bulk copies, one per occupied cell.  Each lane keeps one staged point q in registers and the warp sweeps
// A's own points child by child through broadcast LDS.128: every (p_i, q) pair gets the exact fp32 predicate and a
// lane remembers WHICH of A's children hit it -- no branches inside the sweep, rings 1 and 2 covered at once.
// Lanes are grouped by fine cell (match.any on the fine id), one ballot per child turns lane hits into
// (child a, fine cell b) adjacencies.  The connected components among A's <= 8 children are tracked in a packed
// warp-uniform register (8 x 8-bit masks), every adjacent neighbour cell remembers one child it touches, and at the
// end ONE global edge per touched fine cell goes into the lock-free union-find (atomicMin hooking).
// Tasks whose neighbourhood is too large for the tile are appended to the dense list for k_uf_dense.
__global__ void __launch_bounds__(UFC_THREADS, UFC_MIN_BLOCKS) k_uf_sparse(const float4* __restrict__ spts, const int* __restrict__ fc_start,
                                                            const int4* __restrict__ crec, const int* __restrict__ nbr,
                                                            int* __restrict__ d_counts, int* parent, float r2, int use_tma,
                                                            int* __restrict__ dense_list, int dense_cap) {
    extern __shared__ __align__(16) unsigned char ufc_smem_raw[];
    UfcWarpSmem& sm = reinterpret_cast<UfcWarpSmem*>(ufc_smem_raw)[warp_id()];
    const int lane = lane_id();
    const int n_tasks = d_counts[CNT_COARSE];
    const int n_warps = gridDim.x * UFC_WARPS;
    if (lane == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t parity = 0;
    // software pipeline of the lookup chain: neighbour ids are fetched two tasks ahead, records one task ahead
    const int first = blockIdx.x * UFC_WARPS + warp_id();
    int nci_1 = -1, nci_2 = -1;
    int4 rec_1 = make_int4(0, 0, 0, 0);
    if (lane < 16) {
        if (first < n_tasks) nci_1 = __ldg(nbr + first * 16 + lane);
        if (first + n_warps < n_tasks) nci_2 = __ldg(nbr + (first + n_warps) * 16 + lane);
    }
    if (nci_1 >= 0) rec_1 = __ldg(crec + nci_1);

    for (int ci = first; ci < n_tasks; ci += n_warps) {
        const int4 rec = rec_1;
        rec_1 = make_int4(0, 0, 0, 0);
        if (nci_2 >= 0) rec_1 = __ldg(crec + nci_2);       // record of the next task (consumed next iteration)
        nci_2 = -1;
        if (lane < 16 && ci + 2 * n_warps < n_tasks) nci_2 = __ldg(nbr + (ci + 2 * n_warps) * 16 + lane);
        const int incl = warp_inclusive_scan(rec.y);
        const int ptot = __shfl_sync(kFull, incl, 31);
        const int n_own = __shfl_sync(kFull, rec.y, 0);
        const int own0 = __shfl_sync(kFull, rec.x, 0);
        const int ffirst0 = __shfl_sync(kFull, rec.z, 0);
        const int n_a = __popc((unsigned)__shfl_sync(kFull, rec.w, 0));
        if (!(ptot <= UFC_TILE_PTS && n_own * ptot <= UFC_BRUTE_TESTS)) {
            // dense neighbourhood: hand the task to k_uf_dense
            if (lane == 0) {
                const int slot = atomicAdd(&d_counts[CNT_DENSE], 1);
                if (slot < dense_cap) dense_list[slot] = ci;
                else atomicOr(&d_counts[CNT_FLAGS], 1);  // cannot happen (see dense_cap); the host reports it
            }
            continue;
        }
        if (lane < 16) {
            sm.cstart[lane] = rec.x;
            sm.coff[lane] = incl - rec.y;  // lanes >= 14 hold ptot
            sm.ffirst[lane] = rec.z;
        }
        for (int x = lane; x < UFC_NODES / 4; x += 32) reinterpret_cast<uint32_t*>(sm.attach)[x] = 0xffffffffu;
        // ---- stage the forward neighbourhood: TMA bulk copies (one per occupied cell), or plain loads ----
        bool staged = false;
        if (use_tma & 1) {
            if (lane == 0) mbar_arrive_expect_tx(&sm.bar, (uint32_t)ptot * 16u);
            __syncwarp();
            if (rec.y > 0) tma_load_1d(&sm.tile[incl - rec.y], spts + rec.x, (uint32_t)rec.y * 16u, &sm.bar);
            const bool ok = mbar_wait_bounded(&sm.bar, parity);
            parity ^= 1u;
            staged = __all_sync(kFull, ok);
            if (!staged) {  // the copy never landed: the barrier phase and the tile are in an unknown state -- report, do not guess
                if (lane == 0) atomicOr(&d_counts[CNT_FLAGS], 8);
                return;
            }
        }
        __syncwarp();
        if (!staged) {
            for (int t = lane; t < ptot; t += 32) {
                int c = 0;
#pragma unroll
                for (int k = 8; k > 0; k >>= 1)
                    if (c + k < UFC_CELLS && sm.coff[c + k] <= t) c += k;
                sm.tile[t] = __ldg(spts + sm.cstart[c] + (t - sm.coff[c]));
            }
            __syncwarp();
        }

        // ---- sweep ----
        unsigned long long comp = 0x8040201008040201ull;  // byte a = mask of A's children connected to child a (warp uniform)
        // with a single child nothing inside A needs testing: start at the first chunk that holds a neighbour's point
        int t0 = n_a == 1 ? (n_own & ~31) : 0;
        const unsigned all_children = (1u << n_a) - 1u;
        // General chunks (32 points, one per lane): as long as A's children are not yet known to be one component.
        for (; t0 < ptot && (unsigned)(comp & 0xffull) != all_children; t0 += 32) {
            const int t = t0 + lane;
            const bool valid = t < ptot;
            int c = 0, fid = -1 - lane;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
#pragma unroll
                for (int k = 8; k > 0; k >>= 1)
                    if (c + k < UFC_CELLS && sm.coff[c + k] <= t) c += k;
                q = sm.tile[t];
                fid = __float_as_int(q.w);                    // k_cells_write stores the fine cell id in .w
            }
            unsigned hit = 0;  // bit a: some point of A's child a is within tol of q
            // one flat loop over A's points (p_i is a broadcast LDS.128; its .w carries the child it belongs to, so the
            // child bit is warp-uniform arithmetic) -- long enough trip counts for the unrolled body to be the one that runs
#pragma unroll 4
            for (int i = 0; i < n_own; ++i) {
                const float4 pi = sm.tile[i];
                const unsigned bit = 1u << (unsigned)(__float_as_int(pi.w) - ffirst0);
                if (dist2_exact(pi.x, pi.y, pi.z, q.x, q.y, q.z) < r2) hit |= bit;
            }
            if (!valid) hit = 0;
            const int j = fid - sm.ffirst[c];                 // child index of q's fine cell inside its coarse cell
            if (valid && c == 0) hit &= (1u << j) - 1u;       // inside A each unordered child pair once (a < j)
            const unsigned peers = __match_any_sync(kFull, fid);
            unsigned hm = 0;
            for (int a = 0; a < n_a; ++a) {
                const unsigned bal = __ballot_sync(kFull, (hit >> a) & 1u);
                if (bal & peers) hm |= 1u << a;
            }
            const bool leader = valid && lane == __ffs(peers) - 1 && hm != 0;
            if (!leader) hm = 0;
            if (leader) {
                if (c == 0) hm |= 1u << j;                    // an own child: it joins the children it touches
                else {
                    // a fine cell can straddle two chunks: whatever it touched before is linked to what it touches now
                    const unsigned prev = sm.attach[c * 8 + j];
                    if (prev != 0xffu) hm |= 1u << prev;
                    sm.attach[c * 8 + j] = (unsigned char)(__ffs(hm) - 1);
                }
            }
            // children that this fine cell links for the first time -> merge their components (at most n_a - 1
            // merges per task; the need is re-evaluated after every merge)
            for (;;) {
                const unsigned lowc = hm ? (unsigned)((comp >> (8 * (__ffs(hm) - 1))) & 0xffull) : 0u;
                const unsigned bm = __ballot_sync(kFull, (hm & ~lowc) != 0u);
                if (!bm) break;
                const unsigned m = __shfl_sync(kFull, hm, __ffs(bm) - 1);
                unsigned nc = 0;
                for (unsigned mm = m; mm; mm &= mm - 1) nc |= (unsigned)((comp >> (8 * (__ffs(mm) - 1))) & 0xffull);
                for (unsigned mm = nc; mm; mm &= mm - 1) {
                    const int sh = 8 * (__ffs(mm) - 1);
                    comp = (comp & ~(0xffull << sh)) | ((unsigned long long)nc << sh);
                }
            }
        }
        // Fast chunks (64 points, two per lane): all of A's children are one component, so a lane only has to find out
        // WHETHER its point is within tol of any point of A (a running minimum of the squared distance, nothing else per
        // test); its fine cell then attaches to child 0, which stands for the whole component.  Lanes of the same cell may
        // race, every answer is the same.  Own-cell points (t < n_own) need nothing any more.
        for (; t0 < ptot; t0 += 64) {
            const int ta = t0 + lane, tb = t0 + 32 + lane;
            const float4 qa = ta < ptot ? sm.tile[ta] : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f);  // never within tol of anything
            const float4 qb = tb < ptot ? sm.tile[tb] : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f);
            float da = 3.0e38f, db = 3.0e38f;  // one FMNMX per test, compared once at the end
#pragma unroll 4
            for (int i = 0; i < n_own; ++i) {
                const float4 pi = sm.tile[i];
                da = fminf(da, dist2_exact(pi.x, pi.y, pi.z, qa.x, qa.y, qa.z));
                db = fminf(db, dist2_exact(pi.x, pi.y, pi.z, qb.x, qb.y, qb.z));
            }
            if (da < r2 && ta >= n_own) {
                int c = 0;
#pragma unroll
                for (int k = 8; k > 0; k >>= 1)
                    if (c + k < UFC_CELLS && sm.coff[c + k] <= ta) c += k;
                sm.attach[c * 8 + (__float_as_int(qa.w) - sm.ffirst[c])] = 0;
            }
            if (db < r2 && tb >= n_own) {
                int c = 0;
#pragma unroll
                for (int k = 8; k > 0; k >>= 1)
                    if (c + k < UFC_CELLS && sm.coff[c + k] <= tb) c += k;
                sm.attach[c * 8 + (__float_as_int(qb.w) - sm.ffirst[c])] = 0;
            }
        }
        __syncwarp();
        // ---- one global edge per touched fine cell.  The walks up the global forest are latency bound, so a lane
        // first issues the level-1 and level-2 parent loads of all its (<= 5) edges back to back and only then hooks.
        int ea[5], eb[5];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int x = lane + 32 * r;
            ea[r] = -1; eb[r] = -1;
            if (x < UFC_NODES) {
                const unsigned a = sm.attach[x];
                if (a != 0xffu) {
                    ea[r] = sm.ffirst[x >> 3] + (x & 7);
                    eb[r] = ffirst0 + __ffs((unsigned)((comp >> (8 * a)) & 0xffull)) - 1;
                }
            }
        }
        ea[4] = -1; eb[4] = -1;
        if (lane < n_a) {
            const int rootc = __ffs((unsigned)((comp >> (8 * lane)) & 0xffull)) - 1;
            if (rootc != lane) { ea[4] = ffirst0 + lane; eb[4] = ffirst0 + rootc; }
        }
        // These two look-ahead levels may come from L1 (ld.ca): a stale parent is still an ancestor-or-former-ancestor
        // of the cell, i.e. a member of its set, which is all uf_unite needs; the hooks themselves use L2 (ld.cg/atomics).
#pragma unroll
        for (int r = 0; r < 5; ++r)
            if (ea[r] >= 0) { ea[r] = __ldca(parent + ea[r]); eb[r] = __ldca(parent + eb[r]); }
#pragma unroll
        for (int r = 0; r < 5; ++r)
            if (ea[r] >= 0) { ea[r] = __ldca(parent + ea[r]); eb[r] = __ldca(parent + eb[r]); }
#pragma unroll
        for (int r = 0; r < 5; ++r)
            if (ea[r] >= 0 && ea[r] != eb[r]) {
                // the ancestors stand in for the cells themselves and are roots more often than not: hook straight away;
                // atomicMin on a node that turns out not to be a root is what uf_unite's own retry path handles
                const int hi = max(ea[r], eb[r]), lo = min(ea[r], eb[r]);
                const int old = atomicMin(parent + hi, lo);
                if (old != hi && old != lo) uf_unite(parent, old, lo);
            }
        __syncwarp();
    }
}

// ---- k_uf_sparse2: the same task, HALF a warp per coarse cell ------------------------------------------------------------
// A LiDAR frame's typical coarse cell has ~9 own points and ~50 in its forward neighbourhood: one or two 32-point chunks,
// so a whole warp per task spends most of its instructions on per-task bookkeeping (lookups, staging, cell search, edge
// emission) that a wider warp does not make faster.  Here a warp takes two ADJACENT coarse cells, 16 lanes each: every
// warp-uniform step of the bookkeeping serves two tasks at once (the two neighbour rows are one coalesced 128-byte load),
// and the sweep runs 16-point chunks per task.  Same staging (TMA into a 128-point tile per task), same exact predicate,
// same hooking protocol as k_uf_sparse; a task that does not fit the smaller tile goes to k_uf_dense.
constexpr int UFP_TILE = 128;
#ifndef UFP_BRUTE_V
#define UFP_BRUTE_V 4096
#endif
constexpr int UFP_BRUTE_TESTS = UFP_BRUTE_V;
struct __align__(16) UfpWarpSmem {
    float4 tile[2][UFP_TILE];
    int cstart[2][16];
    int coff[2][16];
    int ffirst[2][16];
    unsigned char attach[2][UFC_NODES];
    unsigned char elist[2][UFC_NODES + 16];  // compacted list of attached neighbour nodes (emission)
    unsigned char cellof[2][UFP_TILE];       // stencil cell (0..13) of every tile position, written once per task
    uint64_t bar;
};

__device__ __forceinline__ int seg16_inclusive_scan(int v) {
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int t = __shfl_up_sync(kFull, v, o, 16);
        if ((lane_id() & 15) >= o) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(UFC_THREADS, UFC_MIN_BLOCKS) k_uf_sparse2(const float4* __restrict__ spts, const int4* __restrict__ crec,
                                                                             const int* __restrict__ nbr, int* __restrict__ d_counts, int* parent,
                                                                             float r2, int use_tma, int* __restrict__ dense_list, int dense_cap) {
    extern __shared__ __align__(16) unsigned char ufc_smem_raw[];
    UfpWarpSmem& sm = reinterpret_cast<UfpWarpSmem*>(ufc_smem_raw)[warp_id()];
    const int lane = lane_id(), hl = lane & 15, half = lane >> 4;
    const unsigned hmask = 0xffffu << (half * 16);
    const int n_tasks = d_counts[CNT_COARSE];
    const int n_pairs = (n_tasks + 1) >> 1;
    const int n_warps = gridDim.x * UFC_WARPS;
    if (lane == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t parity = 0;
    float4* const tile = sm.tile[half];
    unsigned char* const attach = sm.attach[half];
    // software pipeline of the lookup chain: neighbour rows two pairs ahead, records one pair ahead
    const int first = blockIdx.x * UFC_WARPS + warp_id();
    auto row = [&](int pair) { return (pair < n_pairs && 2 * pair + half < n_tasks) ? __ldg(nbr + pair * 32 + lane) : -1; };
    int nci_1 = row(first), nci_2 = row(first + n_warps);
    int4 rec_1 = make_int4(0, 0, 0, 0);
    if (nci_1 >= 0) rec_1 = __ldg(crec + nci_1);

    for (int pair = first; pair < n_pairs; pair += n_warps) {
        int4 rec = rec_1;
        rec_1 = make_int4(0, 0, 0, 0);
        if (nci_2 >= 0) rec_1 = __ldg(crec + nci_2);
        nci_2 = row(pair + 2 * n_warps);
        const int ci = 2 * pair + half;
        int incl = seg16_inclusive_scan(rec.y);
        int ptot = __shfl_sync(kFull, incl, 15, 16);
        int n_own = __shfl_sync(kFull, rec.y, 0, 16);
        const int ffirst0 = __shfl_sync(kFull, rec.z, 0, 16);
        int n_a = __popc((unsigned)__shfl_sync(kFull, rec.w, 0, 16));
        const bool exists = ci < n_tasks;
        const bool fits = ptot <= UFP_TILE && n_own * ptot <= UFP_BRUTE_TESTS;
        if (exists && !fits && hl == 0) {  // dense neighbourhood: hand the task to k_uf_dense
            const int slot = atomicAdd(&d_counts[CNT_DENSE], 1);
            if (slot < dense_cap) dense_list[slot] = ci;
            else atomicOr(&d_counts[CNT_FLAGS], 1);
        }
        if (!exists || !fits) {  // this half sits the pair out
            rec.y = 0;
            incl = 0;
            ptot = 0;
            n_own = 0;
            n_a = 0;
        }
        if (!__any_sync(kFull, ptot > 0)) continue;
        sm.cstart[half][hl] = rec.x;
        sm.coff[half][hl] = incl - rec.y;  // entries 14, 15 hold ptot
        sm.ffirst[half][hl] = rec.z;
        for (int x = lane; x < 2 * UFC_NODES / 4; x += 32) reinterpret_cast<uint32_t*>(sm.attach)[x] = 0xffffffffu;
        for (int k = 0; k < rec.y; ++k) sm.cellof[half][incl - rec.y + k] = (unsigned char)hl;  // lane hl owns stencil cell hl
        // ---- stage both forward neighbourhoods: one TMA bulk copy per occupied cell, one mbarrier for the pair ----
        bool staged = false;
        if (use_tma & 1) {
            const int ptot_other = __shfl_xor_sync(kFull, ptot, 16);
            if (lane == 0) mbar_arrive_expect_tx(&sm.bar, (uint32_t)(ptot + ptot_other) * 16u);
            __syncwarp();
            if (rec.y > 0) tma_load_1d(&tile[incl - rec.y], spts + rec.x, (uint32_t)rec.y * 16u, &sm.bar);
            const bool ok = mbar_wait_bounded(&sm.bar, parity);
            parity ^= 1u;
            staged = __all_sync(kFull, ok);
            if (!staged) {  // see k_uf_sparse
                if (lane == 0) atomicOr(&d_counts[CNT_FLAGS], 8);
                return;
            }
        }
        __syncwarp();
        if (!staged) {
            for (int t = hl; t < ptot; t += 16) {
                int c = 0;
#pragma unroll
                for (int k = 8; k > 0; k >>= 1)
                    if (c + k < UFC_CELLS && sm.coff[half][c + k] <= t) c += k;
                tile[t] = __ldg(spts + sm.cstart[half][c] + (t - sm.coff[half][c]));
            }
            __syncwarp();
        }

        // ---- sweep: 16-point chunks per task while A's children are not yet one component, then 32-point chunks ----
        unsigned long long comp = 0x8040201008040201ull;  // byte a = mask of A's children connected to child a (uniform per half)
        int t0 = n_a == 1 ? (n_own & ~15) : 0;
        const unsigned all_children = (1u << n_a) - 1u;
        for (;;) {
            const bool work = t0 < ptot;
            if (!__any_sync(kFull, work)) break;
            if (work && (unsigned)(comp & 0xffull) != all_children) {
                // general chunk (see k_uf_sparse); every collective below is confined to this half
                const int t = t0 + hl;
                const bool valid = t < ptot;
                int c = 0, fid = -1 - lane;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    c = sm.cellof[half][t];
                    q = tile[t];
                    fid = __float_as_int(q.w);
                }
                unsigned hit = 0;
#pragma unroll 4
                for (int i = 0; i < n_own; ++i) {
                    const float4 pi = tile[i];
                    const unsigned bit = 1u << (unsigned)(__float_as_int(pi.w) - ffirst0);
                    if (dist2_exact(pi.x, pi.y, pi.z, q.x, q.y, q.z) < r2) hit |= bit;
                }
                if (!valid) hit = 0;
                const int j = fid - sm.ffirst[half][c];
                if (valid && c == 0) hit &= (1u << j) - 1u;
                const unsigned peers = __match_any_sync(hmask, fid);
                unsigned hm = __reduce_or_sync(peers, hit);  // children of A touched by any point of this fine cell
                const bool leader = valid && lane == __ffs(peers) - 1 && hm != 0;
                if (!leader) hm = 0;
                if (leader) {
                    if (c == 0) hm |= 1u << j;
                    else {
                        const unsigned prev = attach[c * 8 + j];
                        if (prev != 0xffu) hm |= 1u << prev;
                        attach[c * 8 + j] = (unsigned char)(__ffs(hm) - 1);
                    }
                }
                for (;;) {
                    const unsigned lowc = hm ? (unsigned)((comp >> (8 * (__ffs(hm) - 1))) & 0xffull) : 0u;
                    const unsigned bm = __ballot_sync(hmask, (hm & ~lowc) != 0u);
                    if (!bm) break;
                    const unsigned m = __shfl_sync(hmask, hm, __ffs(bm) - 1);
                    unsigned nc = 0;
                    for (unsigned mm = m; mm; mm &= mm - 1) nc |= (unsigned)((comp >> (8 * (__ffs(mm) - 1))) & 0xffull);
                    for (unsigned mm = nc; mm; mm &= mm - 1) {
                        const int sh = 8 * (__ffs(mm) - 1);
                        comp = (comp & ~(0xffull << sh)) | ((unsigned long long)nc << sh);
                    }
                }
                t0 += 16;
            } else if (work) {
                // fast chunk: children are one component, a lane only needs to know WHETHER its point touches A
                const int ta = t0 + hl, tb = t0 + 16 + hl;
                const float4 qa = ta < ptot ? tile[ta] : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f);
                const float4 qb = tb < ptot ? tile[tb] : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f);
                float da = 3.0e38f, db = 3.0e38f;
#pragma unroll 4
                for (int i = 0; i < n_own; ++i) {
                    const float4 pi = tile[i];
                    da = fminf(da, dist2_exact(pi.x, pi.y, pi.z, qa.x, qa.y, qa.z));
                    db = fminf(db, dist2_exact(pi.x, pi.y, pi.z, qb.x, qb.y, qb.z));
                }
                if (da < r2 && ta >= n_own) {
                    const int c = sm.cellof[half][ta];
                    attach[c * 8 + (__float_as_int(qa.w) - sm.ffirst[half][c])] = 0;
                }
                if (db < r2 && tb >= n_own) {
                    const int c = sm.cellof[half][tb];
                    attach[c * 8 + (__float_as_int(qb.w) - sm.ffirst[half][c])] = 0;
                }
                t0 += 32;
            }
        }
        __syncwarp();
        // ---- edges: compact the attached neighbour nodes of each task (plus the links among its own children, coded as
        // nodes 0..7 of cell 0) into a short list, then hook 16 edges per round and task, look-ahead first ----
        int n_list = 0;
#pragma unroll
        for (int r = 0; r < UFC_NODES / 16; ++r) {
            const int x = hl + 16 * r;
            bool on;
            if (r == 0 && hl < 8) {  // cell 0 = A itself: node hl is child hl, linked to its component's first child
                on = hl < n_a && (__ffs((unsigned)((comp >> (8 * hl)) & 0xffull)) - 1) != hl;
            } else {
                on = attach[x] != 0xffu;
            }
            const unsigned bal = (__ballot_sync(kFull, on) >> (half * 16)) & 0xffffu;
            if (on) sm.elist[half][n_list + __popc(bal & ((1u << hl) - 1u))] = (unsigned char)x;
            n_list += __popc(bal);
        }
        __syncwarp();
        for (int e0 = 0; __any_sync(kFull, e0 < n_list); e0 += 16) {
            int ea = -1, eb = -1;
            if (e0 + hl < n_list) {
                const int x = sm.elist[half][e0 + hl];
                const int a = x < 8 ? x : attach[x];  // own child, or the child the neighbour node touches
                ea = sm.ffirst[half][x >> 3] + (x & 7);
                eb = ffirst0 + __ffs((unsigned)((comp >> (8 * a)) & 0xffull)) - 1;
            }
            // look-ahead through L1 (stale parents are still set members), then hook the ancestors directly
            if (ea >= 0) { ea = __ldca(parent + ea); eb = __ldca(parent + eb); }
            if (ea >= 0) { ea = __ldca(parent + ea); eb = __ldca(parent + eb); }
            if (ea >= 0 && ea != eb) {
                const int hi = max(ea, eb), lo = min(ea, eb);
                const int old = atomicMin(parent + hi, lo);
                if (old != hi && old != lo) uf_unite(parent, old, lo);
            }
        }
        __syncwarp();
    }
}

// Dense tasks (listed by k_uf_sparse).  RING 1: fine-cell pairs at Chebyshev distance <= 1; RING 2 (second launch,
// after a global compress): distance 2.  A pair whose cells already share a global root is skipped without touching
// a point -- on a densely sampled surface a ring-2 pair is always connected through the ring-1 cell between them --
// otherwise the warp searches cooperatively for ONE witness pair (dense neighbours produce it in the first 32x32
// block) and hooks the two cells in the global union-find.
template <int RING>
__global__ void __launch_bounds__(UFC_THREADS) k_uf_dense(const float4* __restrict__ spts, const int* __restrict__ fc_start,
                                                           const int4* __restrict__ crec, const int* __restrict__ nbr,
                                                           const int* __restrict__ d_counts, int* parent, float r2,
                                                           const int* __restrict__ dense_list, int dense_cap) {
    __shared__ int s_ffirst[UFC_WARPS][16];
    __shared__ int s_cmask[UFC_WARPS][16];
    const int lane = lane_id(), w = warp_id();
    const int n_tasks = min(d_counts[CNT_DENSE], dense_cap);
    const int n_warps = gridDim.x * UFC_WARPS;
    // one warp per (dense task, child a of its coarse cell): 8x more parallelism than a warp per task
    for (int work = blockIdx.x * UFC_WARPS + w; work < n_tasks * 8; work += n_warps) {
        const int task = work >> 3, a = work & 7;
        int nci = -1;
        if (lane < UFC_CELLS) nci = __ldg(nbr + dense_list[task] * 16 + lane);
        int4 rec = make_int4(0, 0, 0, 0);
        if (nci >= 0) rec = __ldg(crec + nci);
        __syncwarp();
        if (lane < 16) {
            s_ffirst[w][lane] = rec.z;
            s_cmask[w][lane] = rec.w;
        }
        __syncwarp();
        const unsigned mask0 = (unsigned)s_cmask[w][0];
        const int n_a = __popc(mask0);
        if (a < n_a) {
            const int fid_a = s_ffirst[w][0] + a;
            const int a0 = __ldg(fc_start + fid_a), a1 = __ldg(fc_start + fid_a + 1);
            const int code_a = __fns(mask0, 0, a + 1);
            const int ax = code_a & 1, ay = (code_a >> 1) & 1, az = code_a >> 2;
            for (int nb0 = 0; nb0 < UFC_NODES; nb0 += 32) {
                const int node = nb0 + lane;
                const int c = node >> 3, j = node & 7;
                bool elig = false;
                if (c < UFC_CELLS && j < __popc((unsigned)s_cmask[w][c]) && (c > 0 || j > a)) {
                    const int cs = 13 + c;
                    const int code_b = __fns((unsigned)s_cmask[w][c], 0, j + 1);
                    const int bx = 2 * ((cs % 3) - 1) + (code_b & 1), by = 2 * (((cs / 3) % 3) - 1) + ((code_b >> 1) & 1),
                              bz = 2 * ((cs / 9) - 1) + (code_b >> 2);
                    const int ring = max(max(abs(bx - ax), abs(by - ay)), abs(bz - az));
                    elig = (RING == 1 ? ring <= 1 : ring == 2) && uf_find(parent, s_ffirst[w][c] + j) != uf_find(parent, fid_a);
                }
                unsigned bm = __ballot_sync(kFull, elig);
                bool dirty = false;  // this warp has hooked something since the roots above were compared
                while (bm) {
                    const int sl = __ffs(bm) - 1;
                    bm &= bm - 1;
                    const int nd = nb0 + sl;
                    const int fid_b = s_ffirst[w][nd >> 3] + (nd & 7);
                    if (dirty) {  // an earlier hook of this chunk may have joined the two cells already
                        int same = 0;
                        if (lane == 0) same = uf_find(parent, fid_b) == uf_find(parent, fid_a);
                        same = __shfl_sync(kFull, same, 0);
                        if (same) continue;
                    }
                    const int b0 = __ldg(fc_start + fid_b), b1 = __ldg(fc_start + fid_b + 1);
                    const bool found = coop_witness(spts, a0, a1, b0, b1, r2);
                    if (found) {
                        if (lane == 0) uf_unite(parent, fid_a, fid_b);
                        dirty = true;
                    }
                    __syncwarp();
                }
            }
        }
    }
}

// K5: pointer jumping.  IN_PLACE compresses parent[] between the two hooking phases; the final call writes
// root[] (every fine cell points straight at its component's smallest fine-cell index).
template <bool IN_PLACE>
__global__ void __launch_bounds__(256) k_uf_flatten(int* parent, int* __restrict__ root, const int* __restrict__ d_counts) {
    const int n_fine = d_counts[CNT_FINE];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_fine; c += gridDim.x * blockDim.x) {
        int r = c;
        for (;;) {
            const int p = ld_cg(parent + r);
            if (p == r) break;
            r = p;
        }
        if (IN_PLACE) st_cg(parent + c, r);
        else root[c] = r;
    }
}

}  // namespace mot
