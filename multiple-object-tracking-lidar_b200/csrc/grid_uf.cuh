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
// Points are radix sorted by key, so the fine cells of one coarse cell are contiguous; the open-addressing
// hash maps a coarse key to the index of its first fine cell.  A warp owns one fine cell A: 27 lanes probe
// the 27 coarse neighbours, the (<= 216) candidate fine cells are compacted into a per-warp list, small
// cell pairs are searched by one lane each, large ones cooperatively by the whole warp with shuffles.
#pragma once
#include "common.cuh"

namespace mot {

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
template <typename KT>
__global__ void __launch_bounds__(256) k_cell_keys(const float4* __restrict__ pts, int m, GridCodec g,
                                                    const int* __restrict__ frame_offsets, KT* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float4 p = ld_stream(pts + i);
    int ix = __double2int_rd(__dmul_rn(__dsub_rn((double)p.x, g.minx), g.inv_e));
    int iy = __double2int_rd(__dmul_rn(__dsub_rn((double)p.y, g.miny), g.inv_e));
    int iz = __double2int_rd(__dmul_rn(__dsub_rn((double)p.z, g.minz), g.inv_e));
    ix = min(max(ix, 0), g.nfx - 1);
    iy = min(max(iy, 0), g.nfy - 1);
    iz = min(max(iz, 0), g.nfz - 1);
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

// counts[0][b] = fine-cell heads in block b's chunk, counts[1][b] = coarse-cell heads.  Also gathers the
// sorted SoA point array: spts[j] = (x, y, z, bits(original index)).
template <typename KT>
__global__ void __launch_bounds__(CELL_THREADS) k_cells_count(const KT* __restrict__ skeys, const uint32_t* __restrict__ svals,
                                                               const float4* __restrict__ pts, float4* __restrict__ spts, int m,
                                                               int chunk, int* __restrict__ counts) {
    __shared__ int red[2][CELL_THREADS / 32];
    const int begin = blockIdx.x * chunk, end = min(m, begin + chunk);
    int nf = 0, nc = 0;
    for (int j = begin + threadIdx.x; j < end; j += CELL_THREADS) {
        const KT k = skeys[j];
        const uint32_t o = svals[j];
        float4 p = pts[o];
        p.w = __int_as_float((int)o);
        spts[j] = p;
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

// d_counts layout (ints): [0] fine cells, [1] coarse cells, [2] kept clusters, [3] kept points, [4] flags
enum { CNT_FINE = 0, CNT_COARSE = 1, CNT_K = 2, CNT_TOTAL = 3, CNT_FLAGS = 4, CNT_M = 5, CNT_N = 8 };

template <typename KT>
__global__ void __launch_bounds__(CELL_THREADS) k_cells_write(const KT* __restrict__ skeys, const uint32_t* __restrict__ svals, int m,
                                                               int chunk, const int* __restrict__ counts, int* __restrict__ fc_start,
                                                               int* __restrict__ cc_first, int* __restrict__ pcell,
                                                               int* __restrict__ parent, int* __restrict__ csize, int* __restrict__ cmin,
                                                               int* __restrict__ crank, KT* __restrict__ hkeys, int* __restrict__ hvals,
                                                               unsigned hmask, int hshift, int* __restrict__ d_counts) {
    __shared__ int scratch[36];
    int fbase = block_prefix_of(counts, blockIdx.x, scratch);
    int cbase = block_prefix_of(counts + gridDim.x, blockIdx.x, scratch);
    const int begin = blockIdx.x * chunk, end = min(m, begin + chunk);
    for (int tb = begin; tb < end; tb += CELL_THREADS) {
        const int j = tb + threadIdx.x;
        int fh = 0, ch = 0;
        KT k = 0;
        if (j < end) {
            k = skeys[j];
            if (j == 0) { fh = 1; ch = 1; }
            else {
                const KT kp = skeys[j - 1];
                fh = (k != kp);
                ch = ((k >> 3) != (kp >> 3));
            }
        }
        int total;
        const int packed = (ch << 16) | fh;
        const int excl = block_exclusive_scan(packed, scratch, &total);
        const int fi = fbase + (excl & 0xffff) + fh - 1;  // index of the fine cell containing j
        if (j < end) {
            pcell[j] = fi;
            if (fh) {
                fc_start[fi] = j;
                parent[fi] = fi;
                csize[fi] = 0;
                cmin[fi] = (int)svals[j];  // stable sort: the first point of a cell has its smallest original index
                crank[fi] = -1;
            }
            if (ch) {
                const int ci = cbase + (excl >> 16);
                cc_first[ci] = fi;
                hash_insert<KT>(hkeys, hvals, hmask, hshift, k >> 3, ci);
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
