// cell_uf.cuh -- K4, third generation: connected components decided on CELLS first, points last.
//
// Same contract as grid_uf.cuh (reference call site MOT.cpp:472-488): the partition into connected components of
//     { (i,j) : fp32 ((dx*dx)+dy*dy)+dz*dz < r2 }.
// k_uf_sparse2 swept every (own point, neighbourhood point) pair of a coarse cell by brute force: ~820 warp instructions
// per coarse cell, issue bound at 4 % of the HBM roofline.  Here the work moves from points to cells:
//
//   k_cell_local   one thread per coarse cell: the tight fp32 AABB of each of its (<= 8) fine cells and of the coarse cell,
//                  then the connected components AMONG its own fine cells in registers (no atomics); parent[] of a fine
//                  cell starts at its local root, the packed labels go into the coarse record.
//   k_uf_cross     one thread per (coarse cell A, forward neighbour row): finds the <= 3 neighbour cells of the row (one
//                  hash probe + adjacency in the sorted coarse keys), skips a pair whose local roots already share a global
//                  root, rejects it from the coarse boxes, and otherwise walks the fine-cell pairs within Chebyshev
//                  distance 2.  A fine pair is decided EXACTLY from the two boxes whenever possible:
//                      upper = d2(farthest corners) <  r2  =>  every point pair qualifies      => union, no point touched
//                      lower = d2(nearest gap)      >= r2  =>  no point pair can qualify       => skip
//                  (the fp32 predicate is monotone in |dx|, |dy|, |dz| under round-to-nearest, and fl(a-b) is monotone in
//                  a and -b, so both bounds hold for the rounded arithmetic, not just the real one).  Only ambiguous pairs
//                  look at points: a short serial witness search with early exit when the two cells are small, else the
//                  pair goes to a list for k_uf_heavy.  ONE global edge per connected (local component, local component)
//                  pair goes into the lock-free union-find; a thread stops as soon as everything it could connect is
//                  connected.
//   k_uf_heavy     one warp per listed fine-cell pair (ring 1 first, ring 2 after a pointer-jumping pass): root comparison,
//                  then a cooperative witness search pruned by the two boxes.
#pragma once
#include "common.cuh"
#include "grid_uf.cuh"

namespace mot {

enum { CNT_HEAVY1 = 9, CNT_HEAVY2 = 10, CNT_UNITES = 11, CNT_SERIAL_FALLBACK = 12 };

constexpr int CL_SERIAL_MAX = 32;  // fine cells with more points get their box from the whole warp

// ---- diagnostics (build with -DMOT_UF_STATS; read by mot_debug_stats) ---------------------------------------------------
enum { ST_FINDS, ST_HOPS, ST_UNITES, ST_CAS_RETRY, ST_FINE_PAIRS, ST_WITNESS_TESTS, ST_CROSS_PAIRS, ST_ROOT_SKIPS, ST_CBOX_REJECTS, ST_ACCEPTS,
       ST_REJECTS, ST_LOCAL_PAIRS, ST_MAXHOPS, ST_N = 16 };
#ifdef MOT_UF_STATS
__device__ unsigned long long g_uf_stats[ST_N];
#define UFSTAT(i, v) atomicAdd(&g_uf_stats[i], (unsigned long long)(v))
#define UFSTAT_MAX(i, v) atomicMax(&g_uf_stats[i], (unsigned long long)(v))
#else
#define UFSTAT(i, v) ((void)0)
#define UFSTAT_MAX(i, v) ((void)0)
#endif

// ---- self checks (build with -DMOT_CHECKS): compute-sanitizer is closed on the GPU pool, so the index arithmetic of this file
// can be checked by a build of its own -- every violated bound raises flag 16 (the host then fails the call with MOT_ERR_CUDA) ----
#ifdef MOT_CHECKS
#define MOT_CHECK(cond, d_counts) do { if (!(cond)) atomicOr((d_counts) + CNT_FLAGS, 16); } while (0)
#else
#define MOT_CHECK(cond, d_counts) ((void)0)
#endif

// ---- union-find primitives of this generation -----------------------------------------------------------------------------
// MOT_UF_HOOK = 0: larger index under smaller (atomicMin), as grid_uf.cuh.  = 1 (default): the root whose BIT-REVERSED index
// is larger goes under the other (atomicCAS on a root).  Hooking by plain index turns a run of x-adjacent cells -- all united
// at the same moment by neighbouring threads -- into one parent chain as long as the run; with the bit-reversed order a
// contiguous index range becomes a balanced tree (the winner of a range is its multiple of the largest power of two), so a
// find walks O(log run) parents.  Which node ends up as the root is irrelevant downstream: labels come from cmin[root].
#ifndef MOT_UF_HOOK
#define MOT_UF_HOOK 1
#endif
__device__ __forceinline__ int ufp_find(int* parent, int x) {
    int hops = 0;
    for (;;) {
        const int p = ld_cg(parent + x);
        if (p == x) break;
        const int gp = ld_cg(parent + p);
        if (gp == p) { x = p; ++hops; break; }
        st_cg(parent + x, gp);  // path halving
        x = gp;
        hops += 2;
    }
    UFSTAT(ST_FINDS, 1);
    UFSTAT(ST_HOPS, hops);
    UFSTAT_MAX(ST_MAXHOPS, hops);
    return x;
}
// find that starts from an already loaded parent of x (lets the caller issue several first hops together)
__device__ __forceinline__ int ufp_find_from(int* parent, int x, int p) {
    if (p == x) { UFSTAT(ST_FINDS, 1); return x; }
    const int gp = ld_cg(parent + p);
    if (gp == p) { UFSTAT(ST_FINDS, 1); UFSTAT(ST_HOPS, 1); return p; }
    st_cg(parent + x, gp);
    return ufp_find(parent, gp);
}
__device__ __forceinline__ bool ufp_before(int a, int b) {  // a wins (stays root) against b
#if MOT_UF_HOOK == 1
    return __brev((unsigned)a) < __brev((unsigned)b);
#else
    return a < b;
#endif
}
__device__ __forceinline__ void ufp_unite(int* parent, int a, int b) {
    UFSTAT(ST_UNITES, 1);
    for (;;) {
        a = ufp_find(parent, a);
        b = ufp_find(parent, b);
        if (a == b) return;
        if (ufp_before(a, b)) { const int t = a; a = b; b = t; }  // a = loser, hooked under b
#if MOT_UF_HOOK == 1
        const int old = atomicCAS(parent + a, a, b);
#else
        const int old = atomicMin(parent + a, b);
#endif
        if (old == a) return;  // a was still a root: done
        UFSTAT(ST_CAS_RETRY, 1);
        a = old;               // a had been hooked meanwhile: connect where it went with b
    }
}

// per-axis pieces of the two bounds for boxes [alo, ahi] and [blo, bhi]
__device__ __forceinline__ void axis_bounds(float alo, float ahi, float blo, float bhi, float& gap, float& far) {
    const float d1 = __fsub_rn(alo, bhi);  // > 0 iff a lies entirely above b
    const float d2 = __fsub_rn(ahi, blo);  // < 0 iff a lies entirely below b
    gap = fmaxf(0.0f, fmaxf(d1, -d2));
    far = fmaxf(fabsf(d1), fabsf(d2));
}
__device__ __forceinline__ float sq3(float x, float y, float z) {  // the reference's association order
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
// lower / upper bound of the fp32 predicate value over all point pairs of two boxes
__device__ __forceinline__ void box_bounds(const float4& alo, const float4& ahi, const float4& blo, const float4& bhi, float& lower, float& upper) {
    float gx, gy, gz, fx, fy, fz;
    axis_bounds(alo.x, ahi.x, blo.x, bhi.x, gx, fx);
    axis_bounds(alo.y, ahi.y, blo.y, bhi.y, gy, fy);
    axis_bounds(alo.z, ahi.z, blo.z, bhi.z, gz, fz);
    lower = sq3(gx, gy, gz);
    upper = sq3(fx, fy, fz);
}
__device__ __forceinline__ float box_lower(const float4& alo, const float4& ahi, const float4& blo, const float4& bhi) {
    float gx, gy, gz, f;
    axis_bounds(alo.x, ahi.x, blo.x, bhi.x, gx, f);
    axis_bounds(alo.y, ahi.y, blo.y, bhi.y, gy, f);
    axis_bounds(alo.z, ahi.z, blo.z, bhi.z, gz, f);
    return sq3(gx, gy, gz);
}
// lower bound of the predicate value between point p and any point of the box
__device__ __forceinline__ float pt_box_lower(const float4& p, const float4& lo, const float4& hi) {
    const float gx = fmaxf(0.0f, fmaxf(__fsub_rn(lo.x, p.x), __fsub_rn(p.x, hi.x)));
    const float gy = fmaxf(0.0f, fmaxf(__fsub_rn(lo.y, p.y), __fsub_rn(p.y, hi.y)));
    const float gz = fmaxf(0.0f, fmaxf(__fsub_rn(lo.z, p.z), __fsub_rn(p.z, hi.z)));
    return sq3(gx, gy, gz);
}

// serial witness search between two small fine cells (points [a0, a0+na) and [b0, b0+nb)), early exit
__device__ __forceinline__ bool light_witness(const float4* __restrict__ spts, int a0, int na, int b0, int nb, const float4& blo, const float4& bhi,
                                              float r2) {
    {   // the first pair decides most searches: request both points at once (one round trip instead of two)
        const float4 p = __ldg(spts + a0), q = __ldg(spts + b0);
        UFSTAT(ST_WITNESS_TESTS, 1);
        if (dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < r2) return true;
    }
    for (int i = 0; i < na; ++i) {
        const float4 p = __ldg(spts + a0 + i);
        if (!(pt_box_lower(p, blo, bhi) < r2)) continue;
        for (int j = i == 0 ? 1 : 0; j < nb; ++j) {
            const float4 q = __ldg(spts + b0 + j);
            UFSTAT(ST_WITNESS_TESTS, 1);
            if (dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < r2) return true;
        }
    }
    return false;
}

__device__ __forceinline__ void heavy_push(int2* __restrict__ list, int cap, int* counter, int fa, int fb, bool& stored) {
    const int slot = atomicAdd(counter, 1);
    stored = slot < cap;
    if (stored) list[slot] = make_int2(fa, fb);
}

// exact decision for one fine-cell pair; returns true if the two cells are adjacent (some point pair qualifies).
// Ambiguous pairs too large for a serial search go to the heavy list (pending = true, result false).
__device__ __forceinline__ bool fine_pair(const float4* __restrict__ spts, const float4& alo, const float4& ahi, const float4& blo, const float4& bhi,
                                          int fa, int fb, int ring, float r2, int light, int2* __restrict__ heavy1, int2* __restrict__ heavy2,
                                          int heavy_cap, int* __restrict__ d_counts) {
    float lower, upper;
    box_bounds(alo, ahi, blo, bhi, lower, upper);
    UFSTAT(ST_FINE_PAIRS, 1);
    if (!(lower < r2)) { UFSTAT(ST_REJECTS, 1); return false; }
    if (upper < r2) { UFSTAT(ST_ACCEPTS, 1); return true; }
    const int a0 = __float_as_int(alo.w), na = __float_as_int(ahi.w);
    const int b0 = __float_as_int(blo.w), nb = __float_as_int(bhi.w);
    if (na <= light && nb <= light && na * nb <= light) return light_witness(spts, a0, na, b0, nb, blo, bhi, r2);
    bool stored;
    if (ring <= 1) heavy_push(heavy1, heavy_cap, d_counts + CNT_HEAVY1, fa, fb, stored);
    else heavy_push(heavy2, heavy_cap, d_counts + CNT_HEAVY2, fa, fb, stored);
    if (stored) return false;
    // list full (cannot happen at the capacities mot_create chooses unless nearly every cell pair is heavy): stay exact, search here
    atomicAdd(d_counts + CNT_SERIAL_FALLBACK, 1);
    return light_witness(spts, a0, na, b0, nb, blo, bhi, r2);
}

// ---- k_cell_local ---------------------------------------------------------------------------------------------------------
// Record of a coarse cell: x = first sorted point, y = point count, z = first fine cell, w = child mask (bits 0..7, child
// code = fz<<2 | fy<<1 | fx) | local-root rank of child rank k at bits 8+3k (3 bits each).  Fine children are consecutive
// fine ids in child-code order.  Box of fine cell f: fbox[2f] = (min xyz, bits(first point)), fbox[2f+1] = (max xyz,
// bits(point count)); cbox likewise per coarse cell (w unused).
//
// A warp takes 32 consecutive coarse cells = one contiguous range of sorted points and at most 256 fine cells.  The points
// are read ONCE, coalesced; lanes are grouped by fine cell (match.any on the id carried in .w), six redux.sync give the
// group's min / max and the group leader folds them into the cell's box in shared memory -- a cell of any size costs the
// same per point.  Then every lane resolves the components among the children of its own coarse cell from those boxes.
constexpr int CLOC_WARPS = 4;
constexpr int CLOC_THREADS = CLOC_WARPS * 32;
constexpr int CLOC_FINE = 256;
constexpr int CLOC_UNROLL = 4;  // 32-point groups loaded back to back before the first is reduced
struct ClocWarpSmem {
    float mn[3][CLOC_FINE];
    float mx[3][CLOC_FINE];
    int start[CLOC_FINE + 4];
    unsigned char code[CLOC_FINE];
};

// boxes of the batch are complete in sm: write them out and let every lane resolve its own coarse cell (warp level)
__device__ __forceinline__ void cloc_finish(ClocWarpSmem& sm, int lane, bool valid, int ci, int f0, int f1, int F0, int nf,
                                            const float4* __restrict__ spts, int* __restrict__ d_counts, float r2, int light,
                                            int4* __restrict__ crec, float4* __restrict__ cbox, float4* __restrict__ fbox, int* __restrict__ parent,
                                            int2* __restrict__ heavy1, int2* __restrict__ heavy2, int heavy_cap) {
    for (int x = lane; x < nf; x += 32) {
        const int st = sm.start[x];
        fbox[2 * (size_t)(F0 + x)] = make_float4(sm.mn[0][x], sm.mn[1][x], sm.mn[2][x], __int_as_float(st));
        fbox[2 * (size_t)(F0 + x) + 1] = make_float4(sm.mx[0][x], sm.mx[1][x], sm.mx[2][x], __int_as_float(sm.start[x + 1] - st));
    }
    if (!valid) return;
    const int l0 = f0 - F0, n_a = f1 - f0;
    auto lo_of = [&](int x) { return make_float4(sm.mn[0][x], sm.mn[1][x], sm.mn[2][x], __int_as_float(sm.start[x])); };
    auto hi_of = [&](int x) { return make_float4(sm.mx[0][x], sm.mx[1][x], sm.mx[2][x], __int_as_float(sm.start[x + 1] - sm.start[x])); };
    unsigned mask = 0;
    float c_mn0 = INFINITY, c_mn1 = INFINITY, c_mn2 = INFINITY, c_mx0 = -INFINITY, c_mx1 = -INFINITY, c_mx2 = -INFINITY;
    for (int k = 0; k < n_a; ++k) {
        const int x = l0 + k;
        mask |= 1u << sm.code[x];
        c_mn0 = fminf(c_mn0, sm.mn[0][x]); c_mn1 = fminf(c_mn1, sm.mn[1][x]); c_mn2 = fminf(c_mn2, sm.mn[2][x]);
        c_mx0 = fmaxf(c_mx0, sm.mx[0][x]); c_mx1 = fmaxf(c_mx1, sm.mx[1][x]); c_mx2 = fmaxf(c_mx2, sm.mx[2][x]);
    }
    if (cbox) {  // only the lock-step variant k_uf_cross rejects cell pairs from the coarse boxes
        cbox[2 * (size_t)ci] = make_float4(c_mn0, c_mn1, c_mn2, 0.0f);
        cbox[2 * (size_t)ci + 1] = make_float4(c_mx0, c_mx1, c_mx2, 0.0f);
    }
    // connected components among the children (all of them are ring-1 neighbours of each other).  Pass 1 merges what the boxes
    // alone prove (no point is read); pass 2 decides the pairs that are still in different components -- most ambiguous pairs
    // have been connected through a third child by then and never reach the witness search.
    unsigned lab = 0x76543210u;  // 4 bits per child rank: smallest rank of its component
    auto relabel = [&](unsigned hi, unsigned lo) {  // every nibble equal to hi becomes lo (nibble-parallel)
        const unsigned x = lab ^ (hi * 0x11111111u);
        const unsigned nz = (((x & 0x77777777u) + 0x77777777u) | x) & 0x88888888u;  // bit 3 of a nibble set <=> nibble != 0
        const unsigned sel = ((~nz & 0x88888888u) >> 3) * 15u;                          // 0xF where the nibble was hi
        lab = (lab & ~sel) | ((lo * 0x11111111u) & sel);
    };
    for (int pass = 0; pass < 2 && n_a >= 2; ++pass) {
        for (int i = 0; i < n_a - 1; ++i) {
            const float4 alo = lo_of(l0 + i), ahi = hi_of(l0 + i);
            for (int j = i + 1; j < n_a; ++j) {
                const unsigned li = (lab >> (4 * i)) & 15u, lj = (lab >> (4 * j)) & 15u;
                if (li == lj) continue;
                const float4 blo = lo_of(l0 + j), bhi = hi_of(l0 + j);
                bool hit;
                if (pass == 0) {
                    float lower, upper;
                    box_bounds(alo, ahi, blo, bhi, lower, upper);
                    hit = upper < r2;
                } else {
                    UFSTAT(ST_LOCAL_PAIRS, 1);
                    hit = fine_pair(spts, alo, ahi, blo, bhi, f0 + i, f0 + j, 1, r2, light, heavy1, heavy2, heavy_cap, d_counts);
                }
                if (hit) relabel(max(li, lj), min(li, lj));
            }
        }
    }
    unsigned lab3 = 0;
    for (int k = 0; k < n_a; ++k) {
        const unsigned l = (lab >> (4 * k)) & 15u;
        lab3 |= l << (3 * k);
        parent[f0 + k] = f0 + (int)l;
    }
    crec[ci] = make_int4(sm.start[l0], sm.start[l0 + n_a] - sm.start[l0], f0, (int)(mask | (lab3 << 8)));
}

// segmented min / max scan over the lanes of one 32-point group (points of a fine cell are consecutive lanes); on return the
// LAST lane of every run holds the run's box
__device__ __forceinline__ void cloc_scan(int lf, float& ax, float& ay, float& az, float& bx, float& by, float& bz) {
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int lo_ = __shfl_up_sync(kFull, lf, o);
        const float tax = __shfl_up_sync(kFull, ax, o), tay = __shfl_up_sync(kFull, ay, o), taz = __shfl_up_sync(kFull, az, o);
        const float tbx = __shfl_up_sync(kFull, bx, o), tby = __shfl_up_sync(kFull, by, o), tbz = __shfl_up_sync(kFull, bz, o);
        if (lane >= o && lo_ == lf) {
            ax = fminf(ax, tax); ay = fminf(ay, tay); az = fminf(az, taz);
            bx = fmaxf(bx, tbx); by = fmaxf(by, tby); bz = fmaxf(bz, tbz);
        }
    }
}

constexpr int CLOC_DENSE_POINTS = 2048;  // a batch with more points goes to k_cell_local_dense (a whole CTA per batch)

__global__ void __launch_bounds__(CLOC_THREADS) k_cell_local(const float4* __restrict__ spts, const int* __restrict__ fc_start,
                                                              const int* __restrict__ cc_first, const unsigned char* __restrict__ fcode,
                                                              int* __restrict__ d_counts, float r2, int light, int4* __restrict__ crec,
                                                              float4* __restrict__ cbox, float4* __restrict__ fbox, int* __restrict__ parent,
                                                              int2* __restrict__ heavy1, int2* __restrict__ heavy2, int heavy_cap,
                                                              int* __restrict__ dense_list, int dense_cap) {
    __shared__ ClocWarpSmem sm_all[CLOC_WARPS];
    ClocWarpSmem& sm = sm_all[warp_id()];
    const int n_coarse = d_counts[CNT_COARSE];
    const int lane = lane_id();
    const int n_warps = gridDim.x * CLOC_WARPS;
    for (int c0 = (blockIdx.x * CLOC_WARPS + warp_id()) * 32; c0 < n_coarse; c0 += n_warps * 32) {
        const int ci = c0 + lane;
        const bool valid = ci < n_coarse;
        const int nvalid = min(32, n_coarse - c0);
        int f0 = 0, f1 = 0;
        if (valid) {
            f0 = __ldg(cc_first + ci);
            f1 = __ldg(cc_first + ci + 1);
        }
        const int F0 = __shfl_sync(kFull, f0, 0), F1 = __shfl_sync(kFull, f1, nvalid - 1);
        const int nf = F1 - F0;  // <= 256
        MOT_CHECK(nf >= 1 && nf <= CLOC_FINE && F0 >= 0 && F1 <= d_counts[CNT_FINE], d_counts);
        // densely sampled surfaces (thousands of points in these 32 cells): one warp walking them is the tail of the launch
        if (dense_list) {
            int pe = 0;
            if (lane == 0) pe = __ldg(fc_start + F1) - __ldg(fc_start + F0);
            pe = __shfl_sync(kFull, pe, 0);
            if (pe > CLOC_DENSE_POINTS) {
                if (lane == 0) {
                    const int slot = atomicAdd(&d_counts[CNT_DENSE], 1);
                    if (slot < dense_cap) dense_list[slot] = c0;
                    else atomicOr(&d_counts[CNT_FLAGS], 1);  // cannot happen: the list holds a batch per 2048 points
                }
                continue;
            }
        }
        for (int x = lane; x < nf; x += 32) {
            sm.mn[0][x] = INFINITY; sm.mn[1][x] = INFINITY; sm.mn[2][x] = INFINITY;
            sm.mx[0][x] = -INFINITY; sm.mx[1][x] = -INFINITY; sm.mx[2][x] = -INFINITY;
            sm.code[x] = fcode[F0 + x];
        }
        for (int x = lane; x <= nf; x += 32) sm.start[x] = __ldg(fc_start + F0 + x);
        __syncwarp();
        const int P0 = sm.start[0], P1 = sm.start[nf];
        // This warp is the only writer of its boxes and a cell has one tail lane per group, so the fold is a plain
        // read-modify-write.
        for (int jb = P0; jb < P1; jb += 32 * CLOC_UNROLL) {
            float4 pts[CLOC_UNROLL];
#pragma unroll
            for (int u = 0; u < CLOC_UNROLL; ++u) {
                const int j = jb + 32 * u + lane;
                pts[u] = j < P1 ? ld_stream(spts + j) : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            }
#pragma unroll
            for (int u = 0; u < CLOC_UNROLL; ++u) {
                if (jb + 32 * u >= P1) break;  // warp uniform
                const int lf = __float_as_int(pts[u].w) - F0;  // lanes past the end: negative, no cell
                MOT_CHECK(jb + 32 * u + lane >= P1 || (lf >= 0 && lf < nf), d_counts);
                float ax = pts[u].x, ay = pts[u].y, az = pts[u].z, bx = ax, by = ay, bz = az;
                cloc_scan(lf, ax, ay, az, bx, by, bz);
                const int lnext = __shfl_down_sync(kFull, lf, 1);
                if (jb + 32 * u + lane < P1 && (lane == 31 || lnext != lf)) {
                    sm.mn[0][lf] = fminf(sm.mn[0][lf], ax); sm.mn[1][lf] = fminf(sm.mn[1][lf], ay); sm.mn[2][lf] = fminf(sm.mn[2][lf], az);
                    sm.mx[0][lf] = fmaxf(sm.mx[0][lf], bx); sm.mx[1][lf] = fmaxf(sm.mx[1][lf], by); sm.mx[2][lf] = fmaxf(sm.mx[2][lf], bz);
                }
                __syncwarp();
            }
        }
        __syncwarp();
        cloc_finish(sm, lane, valid, ci, f0, f1, F0, nf, spts, d_counts, r2, light, crec, cbox, fbox, parent, heavy1, heavy2, heavy_cap);
        __syncwarp();
    }
}

// The batches k_cell_local handed over: one CTA per batch.  All warps sweep the batch's points (same segmented scan; the
// tail lane of a run folds it into the cell's box with shared-memory atomics on order-preserving ints, several warps may
// meet in one cell), then warp 0 finishes the batch exactly as k_cell_local does.
constexpr int CLD_THREADS = 512;
__global__ void __launch_bounds__(CLD_THREADS) k_cell_local_dense(const float4* __restrict__ spts, const int* __restrict__ fc_start,
                                                                   const int* __restrict__ cc_first, const unsigned char* __restrict__ fcode,
                                                                   int* __restrict__ d_counts, float r2, int light, int4* __restrict__ crec,
                                                                   float4* __restrict__ cbox, float4* __restrict__ fbox, int* __restrict__ parent,
                                                                   int2* __restrict__ heavy1, int2* __restrict__ heavy2, int heavy_cap,
                                                                   const int* __restrict__ dense_list, int dense_cap) {
    __shared__ ClocWarpSmem sm;
    __shared__ int imn[3][CLOC_FINE], imx[3][CLOC_FINE];
    __shared__ int s_f[2];
    const int n_coarse = d_counts[CNT_COARSE];
    const int n_list = min(d_counts[CNT_DENSE], dense_cap);
    const int lane = lane_id();
    for (int e = blockIdx.x; e < n_list; e += gridDim.x) {
        const int c0 = dense_list[e];
        const int nvalid = min(32, n_coarse - c0);
        if (threadIdx.x == 0) {
            s_f[0] = __ldg(cc_first + c0);
            s_f[1] = __ldg(cc_first + c0 + nvalid);
        }
        __syncthreads();
        const int F0 = s_f[0], F1 = s_f[1], nf = F1 - F0;
        for (int x = threadIdx.x; x < nf; x += blockDim.x) {
            imn[0][x] = 0x7fffffff; imn[1][x] = 0x7fffffff; imn[2][x] = 0x7fffffff;
            imx[0][x] = (int)0x80000000; imx[1][x] = (int)0x80000000; imx[2][x] = (int)0x80000000;
            sm.code[x] = fcode[F0 + x];
        }
        for (int x = threadIdx.x; x <= nf; x += blockDim.x) sm.start[x] = __ldg(fc_start + F0 + x);
        __syncthreads();
        const int P0 = sm.start[0], P1 = sm.start[nf];
        for (int jb = P0 + warp_id() * 32; jb < P1; jb += blockDim.x) {  // every warp: its 32-point groups, stride = CTA
            const int j = jb + lane;
            const float4 p = j < P1 ? ld_stream(spts + j) : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            const int lf = __float_as_int(p.w) - F0;
            MOT_CHECK(j >= P1 || (lf >= 0 && lf < nf), d_counts);
            float ax = p.x, ay = p.y, az = p.z, bx = ax, by = ay, bz = az;
            cloc_scan(lf, ax, ay, az, bx, by, bz);
            const int lnext = __shfl_down_sync(kFull, lf, 1);
            if (j < P1 && (lane == 31 || lnext != lf)) {
                atomicMin(&imn[0][lf], float_to_ordered(ax)); atomicMin(&imn[1][lf], float_to_ordered(ay)); atomicMin(&imn[2][lf], float_to_ordered(az));
                atomicMax(&imx[0][lf], float_to_ordered(bx)); atomicMax(&imx[1][lf], float_to_ordered(by)); atomicMax(&imx[2][lf], float_to_ordered(bz));
            }
        }
        __syncthreads();
        for (int x = threadIdx.x; x < nf; x += blockDim.x) {
            sm.mn[0][x] = ordered_to_float_bits(imn[0][x]); sm.mn[1][x] = ordered_to_float_bits(imn[1][x]); sm.mn[2][x] = ordered_to_float_bits(imn[2][x]);
            sm.mx[0][x] = ordered_to_float_bits(imx[0][x]); sm.mx[1][x] = ordered_to_float_bits(imx[1][x]); sm.mx[2][x] = ordered_to_float_bits(imx[2][x]);
        }
        __syncthreads();
        if (warp_id() == 0) {
            const int ci = c0 + lane;
            const bool valid = ci < n_coarse;
            int f0 = 0, f1 = 0;
            if (valid) {
                f0 = __ldg(cc_first + ci);
                f1 = __ldg(cc_first + ci + 1);
            }
            cloc_finish(sm, lane, valid, ci, f0, f1, F0, nf, spts, d_counts, r2, light, crec, cbox, fbox, parent, heavy1, heavy2, heavy_cap);
        }
        __syncthreads();
    }
}

// ---- k_uf_cross -----------------------------------------------------------------------------------------------------------
// children of the neighbour (codes) that can lie within Chebyshev distance 2 of a child with bit `abit` on this axis, when the
// neighbour's coarse offset on the axis is d: fine offset = 2 d + b - a must stay in [-2, 2]
__device__ __forceinline__ unsigned axis_allowed(int d, unsigned abit, unsigned m0, unsigned m1) {
    if (d > 0) return abit ? (m0 | m1) : m0;
    if (d < 0) return abit ? m1 : (m0 | m1);
    return m0 | m1;
}

struct CrossCtx {
    const float4* __restrict__ spts;
    const int4* __restrict__ crec;
    const float4* __restrict__ cbox;
    const float4* __restrict__ fbox;
    int* parent;
    int* d_counts;
    int2* heavy1;
    int2* heavy2;
    int heavy_cap;
    int light;
    float r2;
};

// cheap, uniform part of a cell pair: already one component? boxes too far apart?  true = the pair needs the fine-cell walk.
// rb and the first parent hops (pa = parent[ra.z], pb = parent[rb.z]) were loaded by the caller, all neighbours at once.
__device__ __forceinline__ bool cross_precheck(const CrossCtx& cx, const int4& ra, const float4& calo, const float4& cahi, int B, const int4& rb,
                                               int pa, int pb) {
    UFSTAT(ST_CROSS_PAIRS, 1);
    if ((((unsigned)ra.w | (unsigned)rb.w) >> 8) == 0u && ufp_find_from(cx.parent, ra.z, pa) == ufp_find_from(cx.parent, rb.z, pb)) {
        UFSTAT(ST_ROOT_SKIPS, 1);
        return false;
    }
    const float4 cblo = __ldg(cx.cbox + 2 * (size_t)B), cbhi = __ldg(cx.cbox + 2 * (size_t)B + 1);
    if (!(box_lower(calo, cahi, cblo, cbhi) < cx.r2)) {
        UFSTAT(ST_CBOX_REJECTS, 1);
        return false;
    }
    return true;
}

// the fine-cell walk of one surviving cell pair (A, B); dir packs the neighbour's coarse offset
__device__ __forceinline__ void cross_task(const CrossCtx& cx, int A, int B, int dir) {
    const int dx = (dir & 3) - 1, dy = ((dir >> 2) & 3) - 1, dz = ((dir >> 4) & 3) - 1;
    const int4 ra = __ldg(cx.crec + A), rb = __ldg(cx.crec + B);
    const unsigned mA = (unsigned)ra.w & 0xffu, mB = (unsigned)rb.w & 0xffu;
    const unsigned labA = (unsigned)ra.w >> 8, labB = (unsigned)rb.w >> 8;
    const int f0A = ra.z, f0B = rb.z;
    const bool single = (labA | labB) == 0u;  // both cells are one local component each
    const float4 cblo = __ldg(cx.cbox + 2 * (size_t)B), cbhi = __ldg(cx.cbox + 2 * (size_t)B + 1);
    unsigned long long conn = 0, chk = 0;  // bit la*8+lb: components known connected / whose global roots were compared
    unsigned ma = mA;
    for (int i = 0; ma; ++i) {
        const unsigned ca = (unsigned)__ffs(ma) - 1u;
        ma &= ma - 1u;
        unsigned cand = mB & axis_allowed(dx, ca & 1u, 0x55u, 0xAAu) & axis_allowed(dy, (ca >> 1) & 1u, 0x33u, 0xCCu) &
                        axis_allowed(dz, ca >> 2, 0x0Fu, 0xF0u);
        if (!cand) continue;
        const float4 alo = __ldg(cx.fbox + 2 * (size_t)(f0A + i)), ahi = __ldg(cx.fbox + 2 * (size_t)(f0A + i) + 1);
        if (!(box_lower(alo, ahi, cblo, cbhi) < cx.r2)) continue;  // this child cannot reach the neighbour cell at all
        const unsigned la = (labA >> (3 * i)) & 7u;
        while (cand) {
            const unsigned cb = (unsigned)__ffs(cand) - 1u;
            cand &= cand - 1u;
            const int j = __popc(mB & ((1u << cb) - 1u));
            const unsigned lb = (labB >> (3 * j)) & 7u;
            const unsigned long long bit = 1ull << (la * 8u + lb);
            if (conn & bit) continue;
            if (!single && !(chk & bit)) {
                chk |= bit;
                if (ufp_find(cx.parent, f0A + (int)la) == ufp_find(cx.parent, f0B + (int)lb)) {
                    conn |= bit;
                    continue;
                }
            }
            const float4 blo = __ldg(cx.fbox + 2 * (size_t)(f0B + j)), bhi = __ldg(cx.fbox + 2 * (size_t)(f0B + j) + 1);
            const int ox = 2 * dx + (int)(cb & 1u) - (int)(ca & 1u), oy = 2 * dy + (int)((cb >> 1) & 1u) - (int)((ca >> 1) & 1u),
                      oz = 2 * dz + (int)(cb >> 2) - (int)(ca >> 2);
            const int ring = max(max(abs(ox), abs(oy)), abs(oz));
            if (!fine_pair(cx.spts, alo, ahi, blo, bhi, f0A + i, f0B + j, ring, cx.r2, cx.light, cx.heavy1, cx.heavy2, cx.heavy_cap, cx.d_counts))
                continue;
            conn |= bit;
            ufp_unite(cx.parent, f0A + (int)la, f0B + (int)lb);
            if (single) return;
        }
    }
}

constexpr int UFX_THREADS = 256;
constexpr int UFX_WARPS = UFX_THREADS / 32;
#ifndef UFX_MIN_BLOCKS
#define UFX_MIN_BLOCKS 4
#endif
// forward neighbour rows of the half stencil, faces first: row 0 = (+x) in the own row; then (dy, dz) = (+1,0), (0,+1), (-1,+1), (+1,+1)
__constant__ int c_row_dy[5] = {0, 1, 0, -1, 1};
__constant__ int c_row_dz[5] = {0, 0, 1, 1, 1};

// A warp walks 32 consecutive coarse cells per step.  Stage 1 (all lanes): the row's <= 3 neighbour cells.  Stage 2 (lanes
// with a neighbour): root comparison + coarse boxes -- most pairs end here.  Survivors go to a per-warp queue in shared
// memory; whenever it holds 32 the warp runs the fine-cell walk with every lane busy (run inline, the few survivors of each
// step kept the warp at ~3 active lanes: profiles/r02_ncu_uf_cross_v1.txt).
template <typename KT>
__global__ void __launch_bounds__(UFX_THREADS, UFX_MIN_BLOCKS) k_uf_cross(const KT* __restrict__ ckey, const int4* __restrict__ crec, const float4* __restrict__ cbox,
                                                           const float4* __restrict__ fbox, const float4* __restrict__ spts,
                                                           const KT* __restrict__ hkeys, const int* __restrict__ hvals, int* __restrict__ d_counts,
                                                           int* parent, GridCodec g, float r2, int light, int2* __restrict__ heavy1,
                                                           int2* __restrict__ heavy2, int heavy_cap, int row_begin, int row_end) {
    __shared__ int4 s_queue[UFX_WARPS][64];
    int4* q = s_queue[warp_id()];
    int qn = 0;  // warp uniform
    const int lane = lane_id();
    const int n_coarse = d_counts[CNT_COARSE];
    const int hb = d_counts[CNT_HB];
    const unsigned hmask = (1u << hb) - 1u;
    const int hshift = 32 - hb;
    CrossCtx cx{spts, crec, cbox, fbox, parent, d_counts, heavy1, heavy2, heavy_cap, light, r2};
    const int n_warps = gridDim.x * UFX_WARPS;
    for (int row = row_begin; row < row_end; ++row) {
        const int dy = c_row_dy[row], dz = c_row_dz[row];
        for (int base = (blockIdx.x * UFX_WARPS + warp_id()) * 32; base < n_coarse; base += n_warps * 32) {
            const int A = base + lane;
            int nb0 = -1, nb1 = -1, nb2 = -1;  // neighbour cells of this row at dx = 0, -1, +1 (centre first: it is the face neighbour)
            if (A < n_coarse) {
                const KT ck = ckey[A];
                KT t = ck;
                const int cxa = (int)(t & (((KT)1 << g.bx) - 1)); t >>= g.bx;
                const int cya = (int)(t & (((KT)1 << g.by) - 1)); t >>= g.by;
                const int cza = (int)(t & (((KT)1 << g.bz) - 1)); t >>= g.bz;
                const int frame = (int)t;
                if (row == 0) {
                    if (cxa + 1 < g.ncx && A + 1 < n_coarse && ckey[A + 1] == ck + 1) nb2 = A + 1;
                } else {
                    const int ny = cya + dy, nz = cza + dz;
                    if (ny >= 0 && ny < g.ncy && nz < g.ncz) {
                        const KT qc = coarse_compose<KT>(g, frame, cxa, ny, nz);
                        const bool has_l = cxa > 0, has_r = cxa + 1 < g.ncx;
                        const int j = hash_find<KT>(hkeys, hvals, hmask, hshift, qc);
                        if (j >= 0) {  // sorted coarse keys: the x neighbours of an occupied cell sit next to it
                            nb0 = j;
                            if (has_l && j > 0 && ckey[j - 1] == qc - 1) nb1 = j - 1;
                            if (has_r && j + 1 < n_coarse && ckey[j + 1] == qc + 1) nb2 = j + 1;
                        } else {
                            if (has_l) nb1 = hash_find<KT>(hkeys, hvals, hmask, hshift, qc - 1);
                            if (has_r) {
                                if (nb1 >= 0) { if (nb1 + 1 < n_coarse && ckey[nb1 + 1] == qc + 1) nb2 = nb1 + 1; }
                                else nb2 = hash_find<KT>(hkeys, hvals, hmask, hshift, qc + 1);
                            }
                        }
                    }
                }
            }
            // everything stage 2 needs, requested back to back: the records of the (<= 3) neighbours, then the first parent hop
            // of every cell involved (the kernel is latency bound: one dependent L2 round trip instead of four)
            int4 ra = make_int4(0, 0, 0, 0), rb0 = ra, rb1 = ra, rb2 = ra;
            float4 calo = make_float4(0.f, 0.f, 0.f, 0.f), cahi = calo;
            int pa = 0, pb0 = 0, pb1 = 0, pb2 = 0;
            if (nb0 >= 0 || nb1 >= 0 || nb2 >= 0) {
                ra = __ldg(crec + A);
                if (nb0 >= 0) rb0 = __ldg(crec + nb0);
                if (nb1 >= 0) rb1 = __ldg(crec + nb1);
                if (nb2 >= 0) rb2 = __ldg(crec + nb2);
                calo = __ldg(cbox + 2 * (size_t)A);
                cahi = __ldg(cbox + 2 * (size_t)A + 1);
                pa = ld_cg(parent + ra.z);
                if (nb0 >= 0) pb0 = ld_cg(parent + rb0.z);
                if (nb1 >= 0) pb1 = ld_cg(parent + rb1.z);
                if (nb2 >= 0) pb2 = ld_cg(parent + rb2.z);
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int B = s == 0 ? nb0 : (s == 1 ? nb1 : nb2);
                const int dx = s == 0 ? 0 : (s == 1 ? -1 : 1);
                bool alive = B >= 0;
                if (alive) alive = cross_precheck(cx, ra, calo, cahi, B, s == 0 ? rb0 : (s == 1 ? rb1 : rb2), pa, s == 0 ? pb0 : (s == 1 ? pb1 : pb2));
                const unsigned m = __ballot_sync(kFull, alive);
                if (alive) q[qn + __popc(m & lanemask_lt())] = make_int4(A, B, (dx + 1) | ((dy + 1) << 2) | ((dz + 1) << 4), 0);
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    const int4 t = q[qn + lane];
                    __syncwarp();
                    cross_task(cx, t.x, t.y, t.z);
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    if (lane < qn) {
        const int4 t = q[lane];
        cross_task(cx, t.x, t.y, t.z);
    }
}

// ---- k_uf_survivors / k_uf_walk: the same decisions, split by how regular they are ------------------------------------------
// k_uf_cross (above) runs the fine-cell walk of 32 surviving cell pairs in lock step: a pair needs 1.2 fine-cell decisions on
// average, but the one lane in 32 that has to reject a dozen candidates holds the other 31 (profiles/r02_ncu_uf_cross_v2.txt:
// 1.2 ms, 12 active lanes, IPC 0.8 at any occupancy).  One thread per candidate fine pair is no answer either: without the early
// exit the work grows five-fold (26 M box tests, 70 M finds, 2.5 M lost CAS races; measured 2.2 ms).  So:
//   k_uf_survivors  the regular part -- neighbour lookup, records, cell-level root comparison (parents through L1: a stale parent
//                   is still a member of the set, equal stale roots prove one component) -- few registers, every lane busy;
//                   appends the surviving (A, B, direction) to a task list with one atomicAdd per warp.
//   k_uf_walk       the irregular part -- persistent lanes: per step a lane decides ONE candidate fine pair of its task, lanes
//                   whose task is finished take the next one from the list, the warp reconverges after every step.
// Phases (neighbour rows, faces first) are separate launches so that the later rows see the unions of the earlier ones.
enum { CNT_TASKS0 = 13, CNT_TICKET0 = 16 };  // d_counts[13 + phase]: listed tasks, d_counts[16 + phase]: tasks handed out (<= 3 phases)

// candidate children of B (codes) for child code ca of A at coarse offset (dx, dy, dz)
__device__ __forceinline__ unsigned cand_mask(unsigned mB, unsigned ca, int dx, int dy, int dz) {
    return mB & axis_allowed(dx, ca & 1u, 0x55u, 0xAAu) & axis_allowed(dy, (ca >> 1) & 1u, 0x33u, 0xCCu) & axis_allowed(dz, ca >> 2, 0x0Fu, 0xF0u);
}
// root as seen through L1 (no writes): a former or current member of x's set
__device__ __forceinline__ int ufp_root_cached(const int* parent, int x) {
    for (;;) {
        const int p = __ldca(parent + x);
        if (p == x) return x;
        x = p;
    }
}

constexpr int UFS_THREADS = 256;
template <typename KT>
__global__ void __launch_bounds__(UFS_THREADS) k_uf_survivors(const KT* __restrict__ ckey, const int4* __restrict__ crec, const KT* __restrict__ hkeys,
                                                               const int* __restrict__ hvals, int* __restrict__ d_counts, const int* parent,
                                                               GridCodec g, unsigned rows_mask, int2* __restrict__ tasks, int task_cap,
                                                               int* __restrict__ task_count) {
    const int lane = lane_id();
    const int n_coarse = d_counts[CNT_COARSE];
    const int hb = d_counts[CNT_HB];
    const unsigned hmask = (1u << hb) - 1u;
    const int hshift = 32 - hb;
    const int n_warps = gridDim.x * (UFS_THREADS / 32);
    for (int row = 0; row < 5; ++row) {
        if (!((rows_mask >> row) & 1u)) continue;
        const int dy = c_row_dy[row], dz = c_row_dz[row];
        for (int base = (blockIdx.x * (UFS_THREADS / 32) + warp_id()) * 32; base < n_coarse; base += n_warps * 32) {
            const int A = base + lane;
            int nb0 = -1, nb1 = -1, nb2 = -1;  // neighbour cells of this row at dx = 0, -1, +1
            if (A < n_coarse) {
                const KT ck = ckey[A];
                KT t = ck;
                const int cxa = (int)(t & (((KT)1 << g.bx) - 1)); t >>= g.bx;
                const int cya = (int)(t & (((KT)1 << g.by) - 1)); t >>= g.by;
                const int cza = (int)(t & (((KT)1 << g.bz) - 1)); t >>= g.bz;
                const int frame = (int)t;
                if (row == 0) {
                    if (cxa + 1 < g.ncx && A + 1 < n_coarse && ckey[A + 1] == ck + 1) nb2 = A + 1;
                } else {
                    const int ny = cya + dy, nz = cza + dz;
                    if (ny >= 0 && ny < g.ncy && nz < g.ncz) {
                        const KT qc = coarse_compose<KT>(g, frame, cxa, ny, nz);
                        const bool has_l = cxa > 0, has_r = cxa + 1 < g.ncx;
                        const int j = hash_find<KT>(hkeys, hvals, hmask, hshift, qc);
                        if (j >= 0) {  // sorted coarse keys: the x neighbours of an occupied cell sit next to it
                            nb0 = j;
                            if (has_l && j > 0 && ckey[j - 1] == qc - 1) nb1 = j - 1;
                            if (has_r && j + 1 < n_coarse && ckey[j + 1] == qc + 1) nb2 = j + 1;
                        } else {
                            if (has_l) nb1 = hash_find<KT>(hkeys, hvals, hmask, hshift, qc - 1);
                            if (has_r) {
                                if (nb1 >= 0) { if (nb1 + 1 < n_coarse && ckey[nb1 + 1] == qc + 1) nb2 = nb1 + 1; }
                                else nb2 = hash_find<KT>(hkeys, hvals, hmask, hshift, qc + 1);
                            }
                        }
                    }
                }
            }
            // records of the neighbours, requested back to back; then the cell-level root comparison
            unsigned keep = 0;  // bit s: the pair (A, neighbour s) survives
            if (nb0 >= 0 || nb1 >= 0 || nb2 >= 0) {
                const int4 ra = __ldg(crec + A);
                int4 rb0 = ra, rb1 = ra, rb2 = ra;
                if (nb0 >= 0) rb0 = __ldg(crec + nb0);
                if (nb1 >= 0) rb1 = __ldg(crec + nb1);
                if (nb2 >= 0) rb2 = __ldg(crec + nb2);
                const bool singleA = ((unsigned)ra.w >> 8) == 0u;
                const int rootA = singleA ? ufp_root_cached(parent, ra.z) : -1;
                if (nb0 >= 0 && !(singleA && ((unsigned)rb0.w >> 8) == 0u && ufp_root_cached(parent, rb0.z) == rootA)) keep |= 1u;
                if (nb1 >= 0 && !(singleA && ((unsigned)rb1.w >> 8) == 0u && ufp_root_cached(parent, rb1.z) == rootA)) keep |= 2u;
                if (nb2 >= 0 && !(singleA && ((unsigned)rb2.w >> 8) == 0u && ufp_root_cached(parent, rb2.z) == rootA)) keep |= 4u;
                UFSTAT(ST_CROSS_PAIRS, (nb0 >= 0) + (nb1 >= 0) + (nb2 >= 0));
                UFSTAT(ST_ROOT_SKIPS, (nb0 >= 0) + (nb1 >= 0) + (nb2 >= 0) - __popc(keep));
            }
            const int mine = __popc(keep);
            const int incl = warp_inclusive_scan(mine);
            const int total = __shfl_sync(kFull, incl, 31);
            if (total == 0) continue;
            int slot = 0;
            if (lane == 0) slot = atomicAdd(task_count, total);
            slot = __shfl_sync(kFull, slot, 0) + incl - mine;
            // task = (A | code << 27, B), code = neighbour slot (dx = 0, -1, +1) | row << 2.  mot_create sizes the list for every
            // neighbour of every coarse cell of a phase, so it cannot overflow; if it ever did the host reports it (flag 4)
            if (slot + mine > task_cap) { atomicOr(d_counts + CNT_FLAGS, 4); continue; }
            if (keep & 1u) tasks[slot] = make_int2((int)((unsigned)A | ((unsigned)(0 | (row << 2)) << 27)), nb0);
            slot += keep & 1u;
            if (keep & 2u) tasks[slot] = make_int2((int)((unsigned)A | ((unsigned)(1 | (row << 2)) << 27)), nb1);
            slot += (keep >> 1) & 1u;
            if (keep & 4u) tasks[slot] = make_int2((int)((unsigned)A | ((unsigned)(2 | (row << 2)) << 27)), nb2);
        }
    }
}

constexpr int UFW_THREADS = 128;
#ifndef UFW_MIN_BLOCKS
#define UFW_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(UFW_THREADS, UFW_MIN_BLOCKS) k_uf_walk(const int2* __restrict__ tasks, const int* __restrict__ task_count, int task_cap,
                                                                          int* __restrict__ ticket, const int4* __restrict__ crec,
                                                                          const float4* __restrict__ fbox, const float4* __restrict__ spts, int* parent,
                                                                          float r2, int light, int2* __restrict__ heavy1, int2* __restrict__ heavy2,
                                                                          int heavy_cap, int* __restrict__ d_counts) {
    const int n = min(*task_count, task_cap);
    const int lane = lane_id();
    // lane state: one cell pair (A, B), an iterator over A's children x the candidate children of B
    bool live = false, single = false;
    int f0A = 0, f0B = 0, dx = 0, dy = 0, dz = 0, i = 0;
    unsigned mA = 0, mB = 0, labA = 0, labB = 0, ma_rem = 0, cand = 0, ca = 0;
    bool exhausted = false;  // warp uniform
    for (;;) {
        const unsigned idle = __ballot_sync(kFull, !live);
        if (idle && !exhausted) {
            const int want = __popc(idle);
            int start = 0;
            if (lane == 0) start = atomicAdd(ticket, want);
            start = __shfl_sync(kFull, start, 0);
            if (start + want >= n) exhausted = true;
            const int t = start + __popc(idle & lanemask_lt());
            if (!live && t < n) {
                const int2 tk = tasks[t];
                const int code = (int)((unsigned)tk.x >> 27), row = code >> 2, sl = code & 3;
                const int4 ra = __ldg(crec + (tk.x & 0x7ffffff)), rb = __ldg(crec + tk.y);
                f0A = ra.z; f0B = rb.z;
                mA = (unsigned)ra.w & 0xffu; mB = (unsigned)rb.w & 0xffu;
                labA = (unsigned)ra.w >> 8; labB = (unsigned)rb.w >> 8;
                single = (labA | labB) == 0u;
                dx = sl == 0 ? 0 : (sl == 1 ? -1 : 1); dy = c_row_dy[row]; dz = c_row_dz[row];
                ma_rem = mA; cand = 0;
                live = true;
            }
        }
        if (!__any_sync(kFull, live)) break;
        if (live) {
            while (!cand) {  // next child of A that has candidates in B
                if (!ma_rem) { live = false; break; }
                ca = (unsigned)__ffs(ma_rem) - 1u;
                ma_rem &= ma_rem - 1u;
                i = __popc(mA & ((1u << ca) - 1u));
                cand = cand_mask(mB, ca, dx, dy, dz);
            }
        }
        if (live) {
            const unsigned cb = (unsigned)__ffs(cand) - 1u;
            cand &= cand - 1u;
            const int j = __popc(mB & ((1u << cb) - 1u));
            const int fa = f0A + i, fb = f0B + j;
            const int la = f0A + (int)((labA >> (3 * i)) & 7u), lb = f0B + (int)((labB >> (3 * j)) & 7u);  // local roots
            const float4 alo = __ldg(fbox + 2 * (size_t)fa), ahi = __ldg(fbox + 2 * (size_t)fa + 1);
            const float4 blo = __ldg(fbox + 2 * (size_t)fb), bhi = __ldg(fbox + 2 * (size_t)fb + 1);
            int pa = 0, pb = 0;
            if (!single) { pa = ld_cg(parent + la); pb = ld_cg(parent + lb); }
            float lower, upper;
            box_bounds(alo, ahi, blo, bhi, lower, upper);
            UFSTAT(ST_FINE_PAIRS, 1);
            bool todo = lower < r2;
            if (!todo) UFSTAT(ST_REJECTS, 1);
            // several local components: skip what is already one global component (a single-single task was compared when listed)
            if (todo && !single && ufp_find_from(parent, la, pa) == ufp_find_from(parent, lb, pb)) todo = false;
            if (todo) {
                bool hit = upper < r2;
                if (hit) UFSTAT(ST_ACCEPTS, 1);
                if (!hit) {
                    const int a0 = __float_as_int(alo.w), na = __float_as_int(ahi.w);
                    const int b0 = __float_as_int(blo.w), nb = __float_as_int(bhi.w);
                    if (na <= light && nb <= light && na * nb <= light) {
                        hit = light_witness(spts, a0, na, b0, nb, blo, bhi, r2);
                    } else {
                        const int ox = 2 * dx + (int)(cb & 1u) - (int)(ca & 1u), oy = 2 * dy + (int)((cb >> 1) & 1u) - (int)((ca >> 1) & 1u),
                                  oz = 2 * dz + (int)(cb >> 2) - (int)(ca >> 2);
                        bool stored;
                        if (max(max(abs(ox), abs(oy)), abs(oz)) <= 1) heavy_push(heavy1, heavy_cap, d_counts + CNT_HEAVY1, fa, fb, stored);
                        else heavy_push(heavy2, heavy_cap, d_counts + CNT_HEAVY2, fa, fb, stored);
                        if (!stored) {  // list full: stay exact, search here
                            atomicAdd(d_counts + CNT_SERIAL_FALLBACK, 1);
                            hit = light_witness(spts, a0, na, b0, nb, blo, bhi, r2);
                        }
                    }
                }
                if (hit) {
                    ufp_unite(parent, la, lb);
                    if (single) live = false;  // both cells are one component each: nothing left to connect
                }
            }
        }
        __syncwarp();
    }
}

// ---- k_uf_fused: stage 1-2 of k_uf_cross + the persistent-lane walk of k_uf_walk in ONE kernel -------------------------------
// Unions become visible to the root comparison of the very next cells (6.2 M of 8.7 M cell pairs end there, against 3 M when
// the phases are separate launches), and the walk keeps its lanes busy: survivors wait in a per-warp queue in shared memory,
// every lane owns one cell pair at a time and decides one candidate fine pair per step, finished lanes pop the next pair.
struct WalkLane {          // one cell pair being walked, packed (lives in registers across stage 1-2)
    int f0A, f0B;          // first fine cell of A / B
    unsigned labA, labB;   // local-root ranks, 3 bits per child
    unsigned masks;        // mA | mB << 8 | remaining children of A << 16 | candidates of the current child << 24
    unsigned misc;         // bit 0 live, 1 single, 2-3 dx+1, 4-5 dy+1, 6-7 dz+1, 8-10 code of the current child, 11-13 its rank
};

// queue entry: x = A | code << 27 (code = dx+1 | row << 2), y = B -- 8 bytes, so that ten resident CTAs leave the L1 alone
__device__ __forceinline__ void walk_load(WalkLane& w, const int4* __restrict__ crec, const int2& tk) {
    const int code = (int)((unsigned)tk.x >> 27), row = code >> 2;
    const int4 ra = __ldg(crec + (tk.x & 0x7ffffff)), rb = __ldg(crec + tk.y);
    w.f0A = ra.z; w.f0B = rb.z;
    const unsigned mA = (unsigned)ra.w & 0xffu, mB = (unsigned)rb.w & 0xffu;
    w.labA = (unsigned)ra.w >> 8; w.labB = (unsigned)rb.w >> 8;
    w.masks = mA | (mB << 8) | (mA << 16);
    w.misc = 1u | ((w.labA | w.labB) == 0u ? 2u : 0u) | ((unsigned)((code & 3) | ((c_row_dy[row] + 1) << 2) | ((c_row_dz[row] + 1) << 4)) << 2);
}

// one step of a live lane: advance to the next candidate fine pair and decide it
__device__ __forceinline__ void walk_step(WalkLane& w, const float4* __restrict__ fbox, const float4* __restrict__ spts, int* parent, float r2, int light,
                                          int2* __restrict__ heavy1, int2* __restrict__ heavy2, int heavy_cap, int* __restrict__ d_counts) {
    const unsigned mA = w.masks & 0xffu, mB = (w.masks >> 8) & 0xffu;
    const int dx = (int)((w.misc >> 2) & 3u) - 1, dy = (int)((w.misc >> 4) & 3u) - 1, dz = (int)((w.misc >> 6) & 3u) - 1;
    unsigned ma_rem = (w.masks >> 16) & 0xffu, cand = w.masks >> 24;
    unsigned ca = (w.misc >> 8) & 7u, i = (w.misc >> 11) & 7u;
    while (!cand) {  // next child of A that has candidates in B
        if (!ma_rem) { w.misc &= ~1u; return; }
        ca = (unsigned)__ffs(ma_rem) - 1u;
        ma_rem &= ma_rem - 1u;
        i = (unsigned)__popc(mA & ((1u << ca) - 1u));
        cand = cand_mask(mB, ca, dx, dy, dz);
    }
    const unsigned cb = (unsigned)__ffs(cand) - 1u;
    cand &= cand - 1u;
    w.masks = mA | (mB << 8) | (ma_rem << 16) | (cand << 24);
    w.misc = (w.misc & 0xffu) | (ca << 8) | (i << 11);
    const bool single = (w.misc & 2u) != 0u;
    const int j = __popc(mB & ((1u << cb) - 1u));
    const int fa = w.f0A + (int)i, fb = w.f0B + j;
    const int la = w.f0A + (int)((w.labA >> (3 * i)) & 7u), lb = w.f0B + (int)((w.labB >> (3 * j)) & 7u);  // local roots
    MOT_CHECK(fa >= 0 && fb >= 0 && fa < d_counts[CNT_FINE] && fb < d_counts[CNT_FINE] && la <= fa && lb <= fb && la >= w.f0A && lb >= w.f0B, d_counts);
    const float4 alo = __ldg(fbox + 2 * (size_t)fa), ahi = __ldg(fbox + 2 * (size_t)fa + 1);
    const float4 blo = __ldg(fbox + 2 * (size_t)fb), bhi = __ldg(fbox + 2 * (size_t)fb + 1);
    int pa = 0, pb = 0;
    if (!single) { pa = ld_cg(parent + la); pb = ld_cg(parent + lb); }
    float lower, upper;
    box_bounds(alo, ahi, blo, bhi, lower, upper);
    UFSTAT(ST_FINE_PAIRS, 1);
    if (!(lower < r2)) { UFSTAT(ST_REJECTS, 1); return; }
    // several local components: skip what is already one global component (a single-single pair was compared before it was queued)
    if (!single && (pa == pb || ufp_find_from(parent, la, pa) == ufp_find_from(parent, lb, pb))) return;
    bool hit = upper < r2;
    if (hit) UFSTAT(ST_ACCEPTS, 1);
    if (!hit) {
        const int a0 = __float_as_int(alo.w), na = __float_as_int(ahi.w);
        const int b0 = __float_as_int(blo.w), nb = __float_as_int(bhi.w);
        if (na <= light && nb <= light && na * nb <= light) {
            hit = light_witness(spts, a0, na, b0, nb, blo, bhi, r2);
        } else {
            const int ox = 2 * dx + (int)(cb & 1u) - (int)(ca & 1u), oy = 2 * dy + (int)((cb >> 1) & 1u) - (int)((ca >> 1) & 1u),
                      oz = 2 * dz + (int)(cb >> 2) - (int)(ca >> 2);
            bool stored;
            if (max(max(abs(ox), abs(oy)), abs(oz)) <= 1) heavy_push(heavy1, heavy_cap, d_counts + CNT_HEAVY1, fa, fb, stored);
            else heavy_push(heavy2, heavy_cap, d_counts + CNT_HEAVY2, fa, fb, stored);
            if (!stored) {  // list full: stay exact, search here
                atomicAdd(d_counts + CNT_SERIAL_FALLBACK, 1);
                hit = light_witness(spts, a0, na, b0, nb, blo, bhi, r2);
            }
        }
    }
    if (hit) {
        ufp_unite(parent, la, lb);
        if (single) w.misc &= ~1u;  // both cells are one component each: nothing left to connect
    }
}

constexpr int UFF_THREADS = 128;
constexpr int UFF_WARPS = UFF_THREADS / 32;
constexpr int UFF_QUEUE = 128;  // per warp; stage 2 adds at most 96 per step to fewer than 32
#ifndef UFF_MIN_BLOCKS
#define UFF_MIN_BLOCKS 10
#endif
template <typename KT>
__global__ void __launch_bounds__(UFF_THREADS, UFF_MIN_BLOCKS) k_uf_fused(const KT* __restrict__ ckey, const int4* __restrict__ crec,
                                                                           const float4* __restrict__ fbox, const float4* __restrict__ spts,
                                                                           const KT* __restrict__ hkeys, const int* __restrict__ hvals,
                                                                           int* __restrict__ d_counts, int* parent, GridCodec g, float r2, int light,
                                                                           int2* __restrict__ heavy1, int2* __restrict__ heavy2, int heavy_cap,
                                                                           int row_inner, int tile_batches) {
    __shared__ int2 s_queue[UFF_WARPS][UFF_QUEUE];
    int2* q = s_queue[warp_id()];
    int qn = 0;  // warp uniform
    WalkLane w{0, 0, 0u, 0u, 0u, 0u};
    const int lane = lane_id();
    const int n_coarse = d_counts[CNT_COARSE];
    const int hb = d_counts[CNT_HB];
    const unsigned hmask = (1u << hb) - 1u;
    const int hshift = 32 - hb;
    const int n_warps = gridDim.x * UFF_WARPS;
    // walk while the lanes can be kept busy (or, at the end, until nothing is left)
    auto drain = [&](bool all) {
        for (;;) {
            const unsigned idle = __ballot_sync(kFull, !(w.misc & 1u));
            const int n_live = 32 - __popc(idle);
            if (all ? (qn == 0 && n_live == 0) : (qn + n_live < 32)) break;
            const int take = min(qn, __popc(idle));
            const int k = __popc(idle & lanemask_lt());
            MOT_CHECK(qn >= 0 && qn <= UFF_QUEUE, d_counts);
            if (!(w.misc & 1u) && k < take) {
                MOT_CHECK((q[qn - 1 - k].x & 0x7ffffff) < n_coarse && q[qn - 1 - k].y >= 0 && q[qn - 1 - k].y < n_coarse, d_counts);
                walk_load(w, crec, q[qn - 1 - k]);
            }
            qn -= take;
            __syncwarp();
            if (w.misc & 1u) walk_step(w, fbox, spts, parent, r2, light, heavy1, heavy2, heavy_cap, d_counts);
            __syncwarp();
        }
    };
    // work items = (neighbour row, batch of 32 cells), row-major: faces first, and a warp of a small launch gets its share of
    // ALL rows instead of walking the five rows one after the other (that serial chain was the single-frame latency)
    const int n_batches = (n_coarse + 31) >> 5;
    // Item order: tiles of `tile_batches` batches of cells; inside a tile row-major (all cells of the tile for the face rows
    // first, which is what makes the root skips work).  A tile's tables (keys, records, hash slots, parents: ~60 B per cell)
    // stay in L2 across its five row sweeps; with one tile = everything (tile_batches = 0) every sweep streams them from DRAM
    // once the batch is larger than L2.  Frames never share components, so nothing is lost across tile borders but a few
    // root skips at the seam.
    const int tb = tile_batches > 0 ? min(tile_batches, n_batches) : n_batches;
    const int n_tiles = tb > 0 ? (n_batches + tb - 1) / tb : 0;
    for (long long item = blockIdx.x * UFF_WARPS + warp_id(); item < 5ll * tb * n_tiles; item += n_warps) {
        int row, batch;
        if (row_inner) {
            row = (int)(item % 5);
            batch = (int)(item / 5);
        } else {
            const int tile = (int)(item / (5ll * tb));
            const int r = (int)(item - (long long)tile * 5 * tb);
            row = r / tb;
            batch = tile * tb + (r - row * tb);
        }
        if (batch >= n_batches) continue;  // warp uniform
        const int dy = c_row_dy[row], dz = c_row_dz[row];
        {
            const int base = batch * 32;
            const int A = base + lane;
            int nb0 = -1, nb1 = -1, nb2 = -1;  // neighbour cells of this row at dx = 0, -1, +1 (centre first: it is the face neighbour)
            if (A < n_coarse) {
                const KT ck = ckey[A];
                KT t = ck;
                const int cxa = (int)(t & (((KT)1 << g.bx) - 1)); t >>= g.bx;
                const int cya = (int)(t & (((KT)1 << g.by) - 1)); t >>= g.by;
                const int cza = (int)(t & (((KT)1 << g.bz) - 1)); t >>= g.bz;
                const int frame = (int)t;
                if (row == 0) {
                    if (cxa + 1 < g.ncx && A + 1 < n_coarse && ckey[A + 1] == ck + 1) nb2 = A + 1;
                } else {
                    const int ny = cya + dy, nz = cza + dz;
                    if (ny >= 0 && ny < g.ncy && nz < g.ncz) {
                        const KT qc = coarse_compose<KT>(g, frame, cxa, ny, nz);
                        const bool has_l = cxa > 0, has_r = cxa + 1 < g.ncx;
                        const int j = hash_find<KT>(hkeys, hvals, hmask, hshift, qc);
                        if (j >= 0) {  // sorted coarse keys: the x neighbours of an occupied cell sit next to it
                            nb0 = j;
                            if (has_l && j > 0 && ckey[j - 1] == qc - 1) nb1 = j - 1;
                            if (has_r && j + 1 < n_coarse && ckey[j + 1] == qc + 1) nb2 = j + 1;
                        } else {
                            if (has_l) nb1 = hash_find<KT>(hkeys, hvals, hmask, hshift, qc - 1);
                            if (has_r) {
                                if (nb1 >= 0) { if (nb1 + 1 < n_coarse && ckey[nb1 + 1] == qc + 1) nb2 = nb1 + 1; }
                                else nb2 = hash_find<KT>(hkeys, hvals, hmask, hshift, qc + 1);
                            }
                        }
                    }
                }
            }
            // records of the neighbours and the first parent hop of every cell involved, requested back to back
            unsigned keep = 0;
            if (nb0 >= 0 || nb1 >= 0 || nb2 >= 0) {
                const int4 ra = __ldg(crec + A);
                int4 rb0 = ra, rb1 = ra, rb2 = ra;
                if (nb0 >= 0) rb0 = __ldg(crec + nb0);
                if (nb1 >= 0) rb1 = __ldg(crec + nb1);
                if (nb2 >= 0) rb2 = __ldg(crec + nb2);
                const bool singleA = ((unsigned)ra.w >> 8) == 0u;
                int pa = 0, pb0 = 0, pb1 = 0, pb2 = 0;
                if (singleA) {
                    pa = ld_cg(parent + ra.z);
                    if (nb0 >= 0) pb0 = ld_cg(parent + rb0.z);
                    if (nb1 >= 0) pb1 = ld_cg(parent + rb1.z);
                    if (nb2 >= 0) pb2 = ld_cg(parent + rb2.z);
                }
                // equal first hops already prove one component -- without touching the (hot) root of a large component; only the
                // pairs that differ there walk up
                unsigned undecided = 0;
                if (nb0 >= 0) { if (!(singleA && ((unsigned)rb0.w >> 8) == 0u)) keep |= 1u; else if (pb0 != pa) undecided |= 1u; }
                if (nb1 >= 0) { if (!(singleA && ((unsigned)rb1.w >> 8) == 0u)) keep |= 2u; else if (pb1 != pa) undecided |= 2u; }
                if (nb2 >= 0) { if (!(singleA && ((unsigned)rb2.w >> 8) == 0u)) keep |= 4u; else if (pb2 != pa) undecided |= 4u; }
                if (undecided) {
                    const int rootA = ufp_find_from(parent, ra.z, pa);
                    if ((undecided & 1u) && ufp_find_from(parent, rb0.z, pb0) != rootA) keep |= 1u;
                    if ((undecided & 2u) && ufp_find_from(parent, rb1.z, pb1) != rootA) keep |= 2u;
                    if ((undecided & 4u) && ufp_find_from(parent, rb2.z, pb2) != rootA) keep |= 4u;
                }
                UFSTAT(ST_CROSS_PAIRS, (nb0 >= 0) + (nb1 >= 0) + (nb2 >= 0));
                UFSTAT(ST_ROOT_SKIPS, (nb0 >= 0) + (nb1 >= 0) + (nb2 >= 0) - __popc(keep));
            }
            const int mine = __popc(keep);
            const int incl = warp_inclusive_scan(mine);
            int slot = qn + incl - mine;
            if (keep & 1u) q[slot++] = make_int2((int)((unsigned)A | ((unsigned)(1 | (row << 2)) << 27)), nb0);
            if (keep & 2u) q[slot++] = make_int2((int)((unsigned)A | ((unsigned)(0 | (row << 2)) << 27)), nb1);
            if (keep & 4u) q[slot++] = make_int2((int)((unsigned)A | ((unsigned)(2 | (row << 2)) << 27)), nb2);
            qn += __shfl_sync(kFull, incl, 31);
            MOT_CHECK(qn <= UFF_QUEUE, d_counts);
            __syncwarp();
            drain(false);
        }
    }
    drain(true);
}

// ---- k_uf_heavy -----------------------------------------------------------------------------------------------------------
// cooperative witness search between two fine cells, pruned by the boxes: a lane's point takes part only if it can reach
// the other cell's box at all
__device__ __forceinline__ bool coop_witness_boxed(const float4* __restrict__ spts, int a0, int a1, int b0, int b1, const float4& alo,
                                                   const float4& ahi, const float4& blo, const float4& bhi, float r2) {
    const int lane = lane_id();
    {   // probe: 32 scattered (p, q) pairs in one step -- between densely sampled neighbours one of them almost always hits
        const int na = a1 - a0, nb = b1 - b0;
        const float4 p = __ldg(spts + a0 + (int)(((unsigned)lane * 2654435761u >> 8) % (unsigned)na));
        const float4 q = __ldg(spts + b0 + (int)(((unsigned)lane * 40503u + 17u) % (unsigned)nb));
        if (__any_sync(kFull, dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < r2)) return true;
    }
    for (int ia = a0; ia < a1; ia += 32) {
        bool pv = ia + lane < a1;
        const float4 p = pv ? __ldg(spts + ia + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        pv = pv && pt_box_lower(p, blo, bhi) < r2;
        if (!__any_sync(kFull, pv)) continue;
        for (int jb = b0; jb < b1; jb += 32) {
            bool qv = jb + lane < b1;
            const float4 q = qv ? __ldg(spts + jb + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            qv = qv && pt_box_lower(q, alo, ahi) < r2;
            unsigned qm = __ballot_sync(kFull, qv);
            bool hit = false;
            while (qm) {
                const int s = __ffs(qm) - 1;
                qm &= qm - 1;
                const float qx = __shfl_sync(kFull, q.x, s), qy = __shfl_sync(kFull, q.y, s), qz = __shfl_sync(kFull, q.z, s);
                hit |= dist2_exact(p.x, p.y, p.z, qx, qy, qz) < r2;
            }
            if (__any_sync(kFull, hit && pv)) return true;
        }
    }
    return false;
}

constexpr int UFH_THREADS = 256;
__global__ void __launch_bounds__(UFH_THREADS) k_uf_heavy(const float4* __restrict__ spts, const float4* __restrict__ fbox,
                                                           const int2* __restrict__ list, const int* __restrict__ count, int cap, int* parent,
                                                           float r2) {
    const int n = min(*count, cap);
    const int lane = lane_id();
    const int n_warps = gridDim.x * (UFH_THREADS / 32);
    for (int e = blockIdx.x * (UFH_THREADS / 32) + warp_id(); e < n; e += n_warps) {
        const int2 pr = list[e];
        int same = 0;
        if (lane == 0) same = ufp_find(parent, pr.x) == ufp_find(parent, pr.y);
        same = __shfl_sync(kFull, same, 0);
        if (same) continue;
        const float4 alo = __ldg(fbox + 2 * (size_t)pr.x), ahi = __ldg(fbox + 2 * (size_t)pr.x + 1);
        const float4 blo = __ldg(fbox + 2 * (size_t)pr.y), bhi = __ldg(fbox + 2 * (size_t)pr.y + 1);
        const int a0 = __float_as_int(alo.w), b0 = __float_as_int(blo.w);
        const bool found = coop_witness_boxed(spts, a0, a0 + __float_as_int(ahi.w), b0, b0 + __float_as_int(bhi.w), alo, ahi, blo, bhi, r2);
        if (found && lane == 0) ufp_unite(parent, pr.x, pr.y);
        __syncwarp();
    }
}

// in-place pointer jumping between the two heavy rings; skipped (one load per thread) when ring 2 has nothing listed
__global__ void __launch_bounds__(256) k_uf_flatten_if(int* parent, const int* __restrict__ d_counts, const int* __restrict__ gate) {
    if (*gate == 0) return;
    const int n_fine = d_counts[CNT_FINE];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_fine; c += gridDim.x * blockDim.x) {
        int r = c;
        for (;;) {
            const int p = ld_cg(parent + r);
            if (p == r) break;
            r = p;
        }
        st_cg(parent + c, r);
    }
}

}  // namespace mot
