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
    for (int i = 0; i < na; ++i) {
        const float4 p = __ldg(spts + a0 + i);
        if (!(pt_box_lower(p, blo, bhi) < r2)) continue;
        for (int j = 0; j < nb; ++j) {
            const float4 q = __ldg(spts + b0 + j);
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
    if (!(lower < r2)) return false;
    if (upper < r2) return true;
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
constexpr int CLOC_THREADS = 128;

template <typename KT>
__global__ void __launch_bounds__(CLOC_THREADS) k_cell_local(const KT* __restrict__ skeys, const float4* __restrict__ spts,
                                                              const int* __restrict__ fc_start, const int* __restrict__ cc_first,
                                                              int* __restrict__ d_counts, float r2, int light, int4* __restrict__ crec,
                                                              float4* __restrict__ cbox, float4* __restrict__ fbox, int* __restrict__ parent,
                                                              int2* __restrict__ heavy1, int2* __restrict__ heavy2, int heavy_cap) {
    const int n_coarse = d_counts[CNT_COARSE];
    const int lane = lane_id();
    const int stride = gridDim.x * CLOC_THREADS;
    const int rounds = (n_coarse + stride - 1) / stride;
    for (int it = 0; it < rounds; ++it) {  // warp-uniform trip count: the cooperative phase needs every lane
        const int ci = it * stride + blockIdx.x * CLOC_THREADS + threadIdx.x;
        const bool valid = ci < n_coarse;
        int f0 = 0, n_a = 0, p0 = 0, p1 = 0;
        unsigned mask = 0, deferred = 0;
        float cminx = INFINITY, cminy = INFINITY, cminz = INFINITY, cmaxx = -INFINITY, cmaxy = -INFINITY, cmaxz = -INFINITY;
        if (valid) {
            f0 = __ldg(cc_first + ci);
            n_a = __ldg(cc_first + ci + 1) - f0;
            p0 = __ldg(fc_start + f0);
            int s = p0;
            for (int k = 0; k < n_a; ++k) {
                const int e = __ldg(fc_start + f0 + k + 1);
                mask |= 1u << (unsigned)(skeys[s] & 7);
                if (e - s <= CL_SERIAL_MAX) {
                    float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
                    for (int i = s; i < e; ++i) {
                        const float4 p = __ldg(spts + i);
                        mnx = fminf(mnx, p.x); mny = fminf(mny, p.y); mnz = fminf(mnz, p.z);
                        mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y); mxz = fmaxf(mxz, p.z);
                    }
                    fbox[2 * (size_t)(f0 + k)] = make_float4(mnx, mny, mnz, __int_as_float(s));
                    fbox[2 * (size_t)(f0 + k) + 1] = make_float4(mxx, mxy, mxz, __int_as_float(e - s));
                    cminx = fminf(cminx, mnx); cminy = fminf(cminy, mny); cminz = fminf(cminz, mnz);
                    cmaxx = fmaxf(cmaxx, mxx); cmaxy = fmaxf(cmaxy, mxy); cmaxz = fmaxf(cmaxz, mxz);
                } else {
                    deferred |= 1u << k;
                }
                s = e;
            }
            p1 = s;
        }
        // large fine cells: the whole warp reduces one cell at a time
        unsigned pend = __ballot_sync(kFull, deferred != 0);
        while (pend) {
            const int src = __ffs(pend) - 1;
            pend &= pend - 1;
            unsigned df = __shfl_sync(kFull, deferred, src);
            const int bf0 = __shfl_sync(kFull, f0, src);
            while (df) {
                const int k = __ffs(df) - 1;
                df &= df - 1;
                const int s = __ldg(fc_start + bf0 + k), e = __ldg(fc_start + bf0 + k + 1);
                float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
                for (int i = s + lane; i < e; i += 32) {
                    const float4 p = __ldg(spts + i);
                    mnx = fminf(mnx, p.x); mny = fminf(mny, p.y); mnz = fminf(mnz, p.z);
                    mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y); mxz = fmaxf(mxz, p.z);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mnx = fminf(mnx, __shfl_xor_sync(kFull, mnx, o)); mny = fminf(mny, __shfl_xor_sync(kFull, mny, o));
                    mnz = fminf(mnz, __shfl_xor_sync(kFull, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(kFull, mxx, o));
                    mxy = fmaxf(mxy, __shfl_xor_sync(kFull, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(kFull, mxz, o));
                }
                if (lane == src) {
                    fbox[2 * (size_t)(bf0 + k)] = make_float4(mnx, mny, mnz, __int_as_float(s));
                    fbox[2 * (size_t)(bf0 + k) + 1] = make_float4(mxx, mxy, mxz, __int_as_float(e - s));
                    cminx = fminf(cminx, mnx); cminy = fminf(cminy, mny); cminz = fminf(cminz, mnz);
                    cmaxx = fmaxf(cmaxx, mxx); cmaxy = fmaxf(cmaxy, mxy); cmaxz = fmaxf(cmaxz, mxz);
                }
            }
        }
        __syncwarp();
        if (!valid) continue;
        cbox[2 * (size_t)ci] = make_float4(cminx, cminy, cminz, 0.0f);
        cbox[2 * (size_t)ci + 1] = make_float4(cmaxx, cmaxy, cmaxz, 0.0f);
        // connected components among the children (all of them are ring-1 neighbours of each other)
        unsigned lab = 0x76543210u;  // 4 bits per child rank: smallest rank of its component
        if (n_a >= 2) {
            for (int i = 0; i < n_a - 1; ++i) {
                const float4 alo = fbox[2 * (size_t)(f0 + i)], ahi = fbox[2 * (size_t)(f0 + i) + 1];
                for (int j = i + 1; j < n_a; ++j) {
                    const unsigned li = (lab >> (4 * i)) & 15u, lj = (lab >> (4 * j)) & 15u;
                    if (li == lj) continue;
                    const float4 blo = fbox[2 * (size_t)(f0 + j)], bhi = fbox[2 * (size_t)(f0 + j) + 1];
                    if (!fine_pair(spts, alo, ahi, blo, bhi, f0 + i, f0 + j, 1, r2, light, heavy1, heavy2, heavy_cap, d_counts)) continue;
                    const unsigned lo = min(li, lj), hi = max(li, lj);
                    for (int k = 0; k < n_a; ++k)
                        if (((lab >> (4 * k)) & 15u) == hi) lab = (lab & ~(15u << (4 * k))) | (lo << (4 * k));
                }
            }
        }
        unsigned lab3 = 0;
        for (int k = 0; k < n_a; ++k) {
            const unsigned l = (lab >> (4 * k)) & 15u;
            lab3 |= l << (3 * k);
            parent[f0 + k] = f0 + (int)l;
        }
        crec[ci] = make_int4(p0, p1 - p0, f0, (int)(mask | (lab3 << 8)));
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

__device__ __forceinline__ void cross_pair(const CrossCtx& cx, const int4& ra, const float4& calo, const float4& cahi, int B, int dx, int dy, int dz) {
    const int4 rb = __ldg(cx.crec + B);
    const unsigned mA = (unsigned)ra.w & 0xffu, mB = (unsigned)rb.w & 0xffu;
    const unsigned labA = (unsigned)ra.w >> 8, labB = (unsigned)rb.w >> 8;
    const int f0A = ra.z, f0B = rb.z;
    const bool single = (labA | labB) == 0u;  // both cells are one local component each
    if (single && uf_find(cx.parent, f0A) == uf_find(cx.parent, f0B)) return;
    const float4 cblo = __ldg(cx.cbox + 2 * (size_t)B), cbhi = __ldg(cx.cbox + 2 * (size_t)B + 1);
    if (!(box_lower(calo, cahi, cblo, cbhi) < cx.r2)) return;
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
                if (uf_find(cx.parent, f0A + (int)la) == uf_find(cx.parent, f0B + (int)lb)) {
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
            uf_unite(cx.parent, f0A + (int)la, f0B + (int)lb);
            if (single) return;
        }
    }
}

constexpr int UFX_THREADS = 256;
// forward neighbour rows of the half stencil, faces first: row 0 = (+x) in the own row; then (dy, dz) = (+1,0), (0,+1), (-1,+1), (+1,+1)
__constant__ int c_row_dy[5] = {0, 1, 0, -1, 1};
__constant__ int c_row_dz[5] = {0, 0, 1, 1, 1};

template <typename KT>
__global__ void __launch_bounds__(UFX_THREADS) k_uf_cross(const KT* __restrict__ ckey, const int4* __restrict__ crec, const float4* __restrict__ cbox,
                                                           const float4* __restrict__ fbox, const float4* __restrict__ spts,
                                                           const KT* __restrict__ hkeys, const int* __restrict__ hvals, int* __restrict__ d_counts,
                                                           int* parent, GridCodec g, float r2, int light, int2* __restrict__ heavy1,
                                                           int2* __restrict__ heavy2, int heavy_cap, int row_begin, int row_end) {
    const int n_coarse = d_counts[CNT_COARSE];
    const int hb = d_counts[CNT_HB];
    const unsigned hmask = (1u << hb) - 1u;
    const int hshift = 32 - hb;
    CrossCtx cx{spts, crec, cbox, fbox, parent, d_counts, heavy1, heavy2, heavy_cap, light, r2};
    const int stride = gridDim.x * UFX_THREADS;
    for (int row = row_begin; row < row_end; ++row) {
        const int dy = c_row_dy[row], dz = c_row_dz[row];
        for (int A = blockIdx.x * UFX_THREADS + threadIdx.x; A < n_coarse; A += stride) {
            const KT ck = ckey[A];
            KT t = ck;
            const int cxa = (int)(t & (((KT)1 << g.bx) - 1)); t >>= g.bx;
            const int cya = (int)(t & (((KT)1 << g.by) - 1)); t >>= g.by;
            const int cza = (int)(t & (((KT)1 << g.bz) - 1)); t >>= g.bz;
            const int frame = (int)t;
            int nb[3] = {-1, -1, -1};  // neighbour cells of this row at dx = 0, -1, +1 (centre first: it is the face neighbour)
            if (row == 0) {
                if (cxa + 1 < g.ncx && A + 1 < n_coarse && ckey[A + 1] == ck + 1) nb[2] = A + 1;
            } else {
                const int ny = cya + dy, nz = cza + dz;
                if (ny < 0 || ny >= g.ncy || nz >= g.ncz) continue;
                const KT qc = coarse_compose<KT>(g, frame, cxa, ny, nz);
                const bool has_l = cxa > 0, has_r = cxa + 1 < g.ncx;
                const int j = hash_find<KT>(hkeys, hvals, hmask, hshift, qc);
                if (j >= 0) {  // sorted coarse keys: the x neighbours of an occupied cell sit next to it
                    nb[0] = j;
                    if (has_l && j > 0 && ckey[j - 1] == qc - 1) nb[1] = j - 1;
                    if (has_r && j + 1 < n_coarse && ckey[j + 1] == qc + 1) nb[2] = j + 1;
                } else {
                    if (has_l) nb[1] = hash_find<KT>(hkeys, hvals, hmask, hshift, qc - 1);
                    if (has_r) {
                        if (nb[1] >= 0) { if (nb[1] + 1 < n_coarse && ckey[nb[1] + 1] == qc + 1) nb[2] = nb[1] + 1; }
                        else nb[2] = hash_find<KT>(hkeys, hvals, hmask, hshift, qc + 1);
                    }
                }
            }
            if (nb[0] < 0 && nb[1] < 0 && nb[2] < 0) continue;
            const int4 ra = __ldg(crec + A);
            const float4 calo = __ldg(cbox + 2 * (size_t)A), cahi = __ldg(cbox + 2 * (size_t)A + 1);
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {  // one inlined body for the three cells of the row
                const int B = s == 0 ? nb[0] : (s == 1 ? nb[1] : nb[2]);
                if (B >= 0) cross_pair(cx, ra, calo, cahi, B, s == 0 ? 0 : (s == 1 ? -1 : 1), dy, dz);
            }
        }
    }
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
        if (lane == 0) same = uf_find(parent, pr.x) == uf_find(parent, pr.y);
        same = __shfl_sync(kFull, same, 0);
        if (same) continue;
        const float4 alo = __ldg(fbox + 2 * (size_t)pr.x), ahi = __ldg(fbox + 2 * (size_t)pr.x + 1);
        const float4 blo = __ldg(fbox + 2 * (size_t)pr.y), bhi = __ldg(fbox + 2 * (size_t)pr.y + 1);
        const int a0 = __float_as_int(alo.w), b0 = __float_as_int(blo.w);
        const bool found = coop_witness_boxed(spts, a0, a0 + __float_as_int(ahi.w), b0, b0 + __float_as_int(bhi.w), alo, ahi, blo, bhi, r2);
        if (found && lane == 0) uf_unite(parent, pr.x, pr.y);
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
