// frame_small.cuh -- the whole per-frame path of a SMALL frame in ONE launch.
//
// The general path (mot_b200.cu: cluster_core) is built for throughput: ~30 kernels and two host round trips per call.
// A frame the reference tracker actually sees (BASELINE config c1: 65,536 points, ~13 k of them kept by removeStatic,
// MOT.cpp:461-491) is so small that those launches ARE the latency: every kernel runs for a few microseconds and the
// GPU idles in between.  k_frame_small does removeStatic -> Euclidean clustering -> size filter -> CSR (ascending
// indices) -> per-cluster statistics -> getCentroid in one kernel launched as ONE thread-block cluster (8 or 16 CTAs
// x 1024 threads).  The phases are separated by the hardware cluster barrier (~0.2 us, flushes L1) instead of kernel
// boundaries, CTA partial sums are exchanged through distributed shared memory, and the host synchronises once.
//
// The algorithm is the general path's cut down to what a small frame needs (same predicates, same output order):
//   A  removeStatic keep bits (rs_keep, bit-for-bit MOT.cpp:674-702) + ordered compaction (block scan + DSMEM exchange)
//   B  coarse cell of every point (edge h = tol (1 + 2^-10), absolute origin: no bounding box needed), open-addressing
//      hash insert, rank of the point inside its cell from the cell counter
//   C  exclusive scan of the cell counters (in place) -> cell start; sum of squared counts (work bound, see below)
//   D  points placed cell by cell (spts); 13 forward neighbour slots + own end per cell (one hash probe per pair of cells)
//   E  one thread per placed point: FLANN's fp32 predicate d2 < r2 against the later points of its own cell and the
//      points of the 13 forward cells; lock-free union-find on placed positions (smaller position wins)
//   F  flatten, component size and smallest original index
//   G  components inside [min, max] -> cluster list; component label of every point
//   H  rank of every cluster in (size desc, smallest index asc) order -- PCL's output order -- by counting (K <= 4096)
//   I  CSR offsets (scan in shared memory, every CTA redundantly)
//   J  members dropped into their cluster's segment (atomic cursor), K: ordered by counting inside the segment
//   L  farthest pair per (cluster, slab) -> M: statistics, line distance, circumcentre per cluster (one warp each)
// Whatever does not fit the small path -- a cell too crowded for a thread-per-point search, more than 4096 clusters,
// clusters too large for the O(size^2) ordering -- raises FS_FLAG_FALLBACK and the host runs the general path on the
// same input: the small path never returns a result it cannot stand behind.
#pragma once
#include <cooperative_groups.h>

#include "cell_uf.cuh"
#include "cluster_table.cuh"
#include "common.cuh"
#include "grid_uf.cuh"
#include "remove_static.cuh"

namespace mot {
namespace cg = cooperative_groups;

constexpr int FS_THREADS = 1024;
constexpr int FS_MAX_K = 4096;            // clusters the counting rank handles (K^2 comparisons)
constexpr int FS_MAX_ITEMS = 32;          // input points per thread in the compaction phase (keep bits live in one word)
enum { FS_ST_T = 0, FS_ST_LOGT = 1, FS_ST_CELLS = 2, FS_ST_EDGES = 3, FS_ST_N = 8 };
constexpr int FS_MAX_NODES = 49152;       // kept points (= union-find node ids) that fit the shared-memory parent array of k_fs_tables (192 KB)
constexpr int FS_CELL_RANGE = 1 << 20;    // |cell coordinate| below this (21 bits per axis in the hash key)
constexpr unsigned long long FS_EMPTY = ~0ull;
constexpr int FS_FLAG_FALLBACK = 32;      // d_counts[CNT_FLAGS]: the general path has to take this frame
constexpr int FS_FLAG_NONFINITE = 64;
constexpr int FS_PHASES = 32;            // 0..15: phase boundaries; 16..: sub-steps of the shared-memory union-find (diagnostics)
constexpr size_t FS_TABLES_SMEM = (size_t)FS_MAX_NODES * 4;  // dynamic shared memory of k_fs_tables (>= (2 FS_MAX_K + 2) ints)

struct FsArgs {
    const float4* src;        // input cloud (device)
    int n;
    int do_rs;                // 1: removeStatic first (src -> kept), 0: src is clustered as it is
    MapParams mp;
    const uint32_t* bits;     // dilated blocked bitmap (mot_set_map)
    float4* kept;             // compaction target when do_rs
    double inv_e;             // 1 / fine cell edge = 2 / (tol (1 + 2^-10))
    float r2;                 // (float)(tol * tol), FLANN's radius
    int min_size, max_size;
    int with_centroids;
    float stamp;
    int cell_cap;             // a fine cell with more points hands the frame back (bounds one thread's witness search)
    int fp_ctas;              // CTAs of k_fs_farthest's grid (both table kernels derive the slab count from it)
    unsigned long long order_limit;  // bound on the sum of squared kept cluster sizes (phases K and L)
    // workspace (all written inside the kernel: never read through the non-coherent path)
    unsigned long long* hkeys;  // [T]
    int* hstart;                // [T] cell counter, then cell start
    int T, log_t;
    int* pslot;                 // [n] hash slot of the point's cell
    int* prank;                 // [n] rank of the point inside its cell
    float4* spts;               // [n] placed points, .w = original index
    int* scell;                 // [n] start of the placed point's cell
    int* celllist;              // [n] slots of the occupied cells
    float4* fbox;               // [2 n] AABB (lo, hi) of every cell, addressed by cell start
    uint32_t* edges;            // [edge_cap] confirmed cell pairs, lo | hi << 16 (node ids < FS_MAX_NODES <= 65536)
    int edge_cap;
    int* state;                 // [FS_ST_N] T, log T, cells of this launch
    int* parent;                // [n]
    int* root;                  // [n]
    int* csize;                 // [n]
    int* cmin;                  // [n]
    int* crank;                 // [n] cluster rank of a root (or -1)
    int* ksize;                 // [n] cluster list: size, smallest index, root
    int* kmin;
    int* kroot;
    int* ssize;                 // [n] sizes in rank order
    int* cursor;                // [n]
    uint32_t* idx_tmp;          // [n]
    PairCand* cands;            // [max(K, warps)]
    // results
    int* labels;                // [M] smallest original index of the point's component
    int* cl_offsets;            // [K + 1]
    uint32_t* indices;          // [total]
    ClusterStat* stats;         // [K]
    float4* centroids;          // [K]
    int* counts;                // d_counts (CNT_M, CNT_K, CNT_TOTAL, CNT_COARSE, CNT_FLAGS)
    int* host_counts;           // pinned host copy of the counters, written by the last kernel (zero-copy: no copy node, no extra gap)
    unsigned long long* phase_ns;  // [FS_PHASES] %globaltimer at the end of every phase (thread 0 of CTA 0; diagnostics)
};

struct FsExchange {  // one CTA's contribution to a cluster-wide prefix / reduction
    int total;
    int cells;
    unsigned long long squares;
};

__device__ __forceinline__ unsigned long long fs_cell_key(int cx, int cy, int cz) {
    return (unsigned long long)(unsigned)(cx + FS_CELL_RANGE) | ((unsigned long long)(unsigned)(cy + FS_CELL_RANGE) << 21) |
           ((unsigned long long)(unsigned)(cz + FS_CELL_RANGE) << 42);
}
__device__ __forceinline__ int fs_hash(unsigned long long key, int log_t) { return (int)((key * 0x9E3779B97F4A7C15ull) >> (64 - log_t)); }
__device__ __forceinline__ unsigned long long fs_ld_key(unsigned long long* p) { return __ldcg(p); }
// slot of `key` or -1
__device__ __forceinline__ int fs_probe(unsigned long long* hkeys, int T, int log_t, unsigned long long key) {
    int slot = fs_hash(key, log_t);
    for (;;) {
        const unsigned long long k = fs_ld_key(hkeys + slot);
        if (k == key) return slot;
        if (k == FS_EMPTY) return -1;
        slot = (slot + 1) & (T - 1);
    }
}
__device__ __forceinline__ bool fs_cell_of(const float4& p, double inv_h, int& cx, int& cy, int& cz) {
    const double fx = __dmul_rn((double)p.x, inv_h), fy = __dmul_rn((double)p.y, inv_h), fz = __dmul_rn((double)p.z, inv_h);
    const double lim = (double)(FS_CELL_RANGE - 2);
    if (!(fabs(fx) < lim) || !(fabs(fy) < lim) || !(fabs(fz) < lim)) return false;  // also NaN
    cx = __double2int_rd(fx); cy = __double2int_rd(fy); cz = __double2int_rd(fz);
    return true;
}

// union-find on placed positions; the smaller position stays root.  Parents are read through L2 (other SMs hook concurrently).
__device__ __forceinline__ int fs_find(int* parent, int x) {
    for (;;) {
        const int p = ld_cg(parent + x);
        if (p == x) return x;
        const int gp = ld_cg(parent + p);
        if (gp == p) return p;
        st_cg(parent + x, gp);  // path halving
        x = gp;
    }
}
__device__ __forceinline__ int fs_link(int* parent, int a, int b) {  // a, b roots when read; returns the surviving root
    for (;;) {
        if (a == b) return a;
        const int lo = min(a, b), hi = max(a, b);
        const int old = atomicCAS(parent + hi, hi, lo);
        if (old == hi) return lo;
        a = fs_find(parent, old);  // hi had been hooked meanwhile: join where it went
        b = fs_find(parent, lo);
    }
}

__device__ __forceinline__ int fs_block_sum(int v, int* s33) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) s33[warp_id()] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < FS_THREADS / 32; ++w) t += s33[w];
    return t;
}
__device__ __forceinline__ unsigned long long fs_block_sum64(unsigned long long v, unsigned long long* s32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if (lane_id() == 0) s32[warp_id()] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (int w = 0; w < FS_THREADS / 32; ++w) t += s32[w];
    return t;
}

// Every CTA publishes `mine`, the cluster barrier makes it visible, every thread reads all of them through DSMEM.
// Returns the sum over the CTAs ranked before this one in `before` and over all CTAs in `all`.  `slot` must not be reused.
__device__ __forceinline__ void fs_exchange(cg::cluster_group& cl, FsExchange* slot, const FsExchange& mine, FsExchange& before, FsExchange& all) {
    __shared__ FsExchange s_before, s_all;
    if (threadIdx.x == 0) *slot = mine;
    cl.sync();
    // lane r of the first warp reads CTA r's contribution: the <= 16 remote reads are in flight together (a loop over the CTAs in
    // every thread serialised 16 x 3 DSMEM round trips, ~2-3 us per exchange)
    if (threadIdx.x < 32) {
        const unsigned rank = cl.block_rank(), nb = cl.num_blocks();
        const unsigned r = threadIdx.x;
        int t = 0, c = 0;
        unsigned long long q = 0;
        if (r < nb) {
            const FsExchange* remote = cl.map_shared_rank(slot, r);
            t = remote->total; c = remote->cells; q = remote->squares;
        }
        int tb = r < rank ? t : 0, cb = r < rank ? c : 0;
        unsigned long long qb = r < rank ? q : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_xor_sync(kFull, t, o); c += __shfl_xor_sync(kFull, c, o); q += __shfl_xor_sync(kFull, q, o);
            tb += __shfl_xor_sync(kFull, tb, o); cb += __shfl_xor_sync(kFull, cb, o); qb += __shfl_xor_sync(kFull, qb, o);
        }
        if (threadIdx.x == 0) {
            s_all.total = t; s_all.cells = c; s_all.squares = q;
            s_before.total = tb; s_before.cells = cb; s_before.squares = qb;
        }
    }
    __syncthreads();
    before = s_before;
    all = s_all;
    __syncthreads();  // s_before / s_all may be rewritten by the next exchange
}

__device__ __forceinline__ void fs_stamp(const FsArgs& a, int gtid, int phase) {
    if (gtid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.phase_ns[phase] = t;
    }
}

// ======================================================================================================================
// Kernel 1 (one cluster): removeStatic + compaction, fine-cell hash, cell starts, placement.
// ======================================================================================================================
__global__ void __launch_bounds__(FS_THREADS, 1) k_fs_front(const FsArgs* __restrict__ ap) {
    const FsArgs a = *ap;
    cg::cluster_group cl = cg::this_cluster();
    __shared__ int sscan[33];
    __shared__ FsExchange xch[2];
    const int tid = threadIdx.x;
    const int rank = (int)cl.block_rank(), nb = (int)cl.num_blocks();
    const int NT = nb * FS_THREADS;
    const int gtid = rank * FS_THREADS + tid;
    fs_stamp(a, gtid, 0);
    if (gtid < CNT_N && gtid != CNT_M) a.counts[gtid] = 0;  // first use is after the next cluster barrier; CNT_M is written below

    // ---- A: removeStatic + ordered compaction ---------------------------------------------------------------------------
    const float4* cloud = a.src;
    int M = a.n;
    if (a.do_rs) {
        // a warp owns 32 * items consecutive points: round k reads points w0 + 32 k + lane (coalesced), four rounds in flight
        const int per_cta = (a.n + nb - 1) / nb;
        const int items = (per_cta + FS_THREADS - 1) / FS_THREADS;  // <= FS_MAX_ITEMS (host)
        const int ce = min(a.n, (rank + 1) * per_cta);
        const int w0 = rank * per_cta + (tid >> 5) * 32 * items;
        const int lane = tid & 31;
        unsigned keep = 0;  // bit k: the point of round k
        for (int k0 = 0; k0 < items; k0 += 4) {
            float4 p[4];
            bool in[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = w0 + 32 * (k0 + k) + lane;
                in[k] = k0 + k < items && i < ce;
                if (in[k]) p[k] = a.src[i];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (in[k] && rs_keep(p[k], a.mp, a.bits)) keep |= 1u << (k0 + k);
        }
        // position of a kept point: CTAs before, warps before (block scan of the warp totals), rounds before, lanes before
        int warp_total = 0;
        for (int k = 0; k < items; ++k) warp_total += __popc(__ballot_sync(kFull, (keep >> k) & 1u));
        int cta_total;
        const int warp_excl = block_exclusive_scan(lane == 0 ? warp_total : 0, sscan, &cta_total);
        FsExchange mine{cta_total, 0, 0ull}, before, all;
        fs_exchange(cl, &xch[0], mine, before, all);
        int pos = before.total + __shfl_sync(kFull, warp_excl, 0);
        for (int k = 0; k < items; ++k) {
            const unsigned m = __ballot_sync(kFull, (keep >> k) & 1u);
            if ((keep >> k) & 1u) a.kept[pos + __popc(m & lanemask_lt())] = a.src[w0 + 32 * k + lane];
            pos += __popc(m);
        }
        M = all.total;
        cloud = a.kept;
    }
    // the hash is sized for the points that are left: T = 2^lt >= 2 M
    int lt = M > 1 ? 32 - __clz(2 * M - 1) : 1;
    lt = max(10, min(lt, a.log_t));
    if (a.log_t < 10) lt = a.log_t;
    const int T = 1 << lt;
    for (int s = gtid; s < T; s += NT) {
        a.hkeys[s] = FS_EMPTY;
        a.hstart[s] = 0;
    }
    if (gtid == 0) a.counts[CNT_M] = M;
    cl.sync();  // kept cloud complete, hash cleared (and the counters, by the first threads)
    if (M > FS_MAX_NODES) {  // uniform: the shared-memory union-find of k_fs_tables cannot hold the frame
        if (gtid == 0) atomicOr(a.counts + CNT_FLAGS, FS_FLAG_FALLBACK);
        return;              // (no shared memory of a neighbour is read after the barrier above)
    }
    fs_stamp(a, gtid, 1);

    // ---- B: fine cell of every point (edge h / 2: a clique), hash insert, rank inside the cell ----------------------------
    for (int i = gtid; i < M; i += NT) {
        const float4 p = cloud[i];
        int cx, cy, cz;
        a.parent[i] = i;
        a.csize[i] = 0;
        a.cmin[i] = 0x7fffffff;
        a.crank[i] = -1;
        a.cursor[i] = 0;
        if (!fs_cell_of(p, a.inv_e, cx, cy, cz)) {
            const bool finite = (fabsf(p.x) < INFINITY) && (fabsf(p.y) < INFINITY) && (fabsf(p.z) < INFINITY);
            atomicOr(a.counts + CNT_FLAGS, finite ? FS_FLAG_FALLBACK : FS_FLAG_NONFINITE);
            a.pslot[i] = 0;
            a.prank[i] = 0;
            continue;
        }
        const unsigned long long key = fs_cell_key(cx, cy, cz);
        int slot = fs_hash(key, lt);
        for (;;) {
            const unsigned long long old = atomicCAS(a.hkeys + slot, FS_EMPTY, key);
            if (old == FS_EMPTY || old == key) break;
            slot = (slot + 1) & (T - 1);
        }
        a.pslot[i] = slot;
        a.prank[i] = atomicAdd(a.hstart + slot, 1);
    }

    // ---- C: cell starts (exclusive scan of the counters, in place) + list of the occupied slots ---------------------------
    cl.sync();
    fs_stamp(a, gtid, 2);
    {
        const int per = (T + NT - 1) / NT;
        const int s0 = gtid * per;
        int sum = 0, occ = 0;
        bool crowded = false;
        for (int k = 0; k < per; ++k) {
            const int s = s0 + k;
            if (s < T) {
                const int c = a.hstart[s];
                sum += c;
                occ += c > 0;
                crowded |= c > a.cell_cap;
            }
        }
        if (crowded) atomicOr(a.counts + CNT_FLAGS, FS_FLAG_FALLBACK);
        int cta_total, cta_cells;
        const int excl = block_exclusive_scan(sum, sscan, &cta_total);
        const int excl_occ = block_exclusive_scan(occ, sscan, &cta_cells);
        FsExchange mine{cta_total, cta_cells, 0ull}, before, all;
        fs_exchange(cl, &xch[1], mine, before, all);
        if (gtid == 0) {
            a.state[FS_ST_T] = T;
            a.state[FS_ST_LOGT] = lt;
            a.state[FS_ST_CELLS] = all.cells;
            a.state[FS_ST_EDGES] = 0;
            a.counts[CNT_COARSE] = all.cells;
        }
        if (ld_cg(a.counts + CNT_FLAGS) != 0) {  // uniform over the cluster; the later kernels see the flag and return at once
            cl.sync();                           // nobody leaves while a neighbour may still read its shared memory
            return;
        }
        int run = before.total + excl, ci = before.cells + excl_occ;
        for (int k = 0; k < per; ++k) {
            const int s = s0 + k;
            if (s < T) {
                const int c = a.hstart[s];
                a.hstart[s] = run;
                run += c;
                if (c > 0) a.celllist[ci++] = s;
            }
        }
    }

    // ---- D: place the points cell by cell ------------------------------------------------------------------------------
    cl.sync();
    fs_stamp(a, gtid, 3);
    for (int i = gtid; i < M; i += NT) {
        const float4 p = cloud[i];
        const int start = a.hstart[a.pslot[i]];
        const int pos = start + a.prank[i];
        a.spts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        a.scell[pos] = start;
    }
    // ---- D2: bounding box of every cell (decides most cell pairs without a point test) ------------------------------------
    cl.sync();
    {
        const int cells = a.state[FS_ST_CELLS];
        for (int ci = gtid; ci < cells; ci += NT) {
            const int s = a.celllist[ci];
            const int b0 = a.hstart[s], b1 = s + 1 < T ? a.hstart[s + 1] : M;
            float4 lo = a.spts[b0], hi = lo;
            for (int j = b0 + 1; j < b1; ++j) {
                const float4 q = a.spts[j];
                lo.x = fminf(lo.x, q.x); lo.y = fminf(lo.y, q.y); lo.z = fminf(lo.z, q.z);
                hi.x = fmaxf(hi.x, q.x); hi.y = fmaxf(hi.y, q.y); hi.z = fmaxf(hi.z, q.z);
            }
            a.fbox[2 * (size_t)b0] = lo;
            a.fbox[2 * (size_t)b0 + 1] = hi;
        }
    }
    fs_stamp(a, gtid, 4);
}

// ======================================================================================================================
// Kernels 2a-2c (whole GPU): connectivity of the fine cells.  The points of a fine cell are a clique (cell diagonal < tol),
// so the cells are the union-find nodes (id = cell start) and a pair of cells needs ONE witness pair d2 < r2.
//   2a k_fs_edges    one thread per (cell, forward neighbour): hash probe, exact box bounds, first witness.  A confirmed
//                    edge (lo, hi) is appended to a list and hooks hi under lo with ONE atomicMin -- no finds, no retries;
//                    parents only ever decrease, so this leaves a forest that already joins every cell to its best neighbour.
//   2b k_fs_compress one thread per cell: parent[x] = root(x).  Flattening per NODE is what makes the per-EDGE pass cheap
//                    (an ncu capture of the single-kernel version had 45 % of its stall samples in find loops on deep trees).
//   2c k_fs_link     one thread per listed edge, all lanes busy: roots differ -> link (CAS, smaller id wins).
// Every step is a dependent L2 round trip: these phases want memory-level parallelism, which is why they run on all SMs
// instead of inside the cluster.
// ======================================================================================================================
constexpr int FS_PAIR_THREADS = 256;
constexpr int FS_PAIR_CTAS_PER_SM = 6;  // what 40 registers allow: the grid is one resident wave (a thread per (cell, x-row of five neighbours)
                                        // with the probes / ranges / boxes of a row requested together was measured too: 30 us instead of 23 --
                                        // the witness searches of a row then run one after the other; the whole warp on one ambiguous pair at a
                                        // time: 39 us)
__global__ void __launch_bounds__(FS_PAIR_THREADS, FS_PAIR_CTAS_PER_SM) k_fs_edges(const FsArgs* __restrict__ ap) {
    const FsArgs a = *ap;
    const int gtid = blockIdx.x * FS_PAIR_THREADS + threadIdx.x, NT = gridDim.x * FS_PAIR_THREADS;
    fs_stamp(a, gtid, 5);
    if (a.counts[CNT_FLAGS] != 0) return;
    const int T = a.state[FS_ST_T], lt = a.state[FS_ST_LOGT], cells = a.state[FS_ST_CELLS], M = a.counts[CNT_M];
    const long long items = (long long)cells * 64;
    const long long rounds = (items + NT - 1) / NT;
    for (long long it = 0; it < rounds; ++it) {
        const long long item = it * NT + gtid;
        int lo = -1, hi = -1;
        const int d = (int)(item & 63);
        if (item < items && d < 62) {
            const int sA = a.celllist[item >> 6];
            const unsigned long long key = a.hkeys[sA];
            const int L = d + 63;  // forward half of the 5 x 5 x 5 block around the cell (centre = 62)
            const int ox = L % 5 - 2, oy = (L / 5) % 5 - 2, oz = L / 25 - 2;
            const int cx = (int)(key & 0x1fffffu) + ox, cy = (int)((key >> 21) & 0x1fffffu) + oy, cz = (int)((key >> 42) & 0x1fffffu) + oz;  // biased
            const unsigned long long nk = (unsigned long long)(unsigned)cx | ((unsigned long long)(unsigned)cy << 21) | ((unsigned long long)(unsigned)cz << 42);
            const int sB = fs_probe(a.hkeys, T, lt, nk);
            if (sB >= 0) {
                const int a0 = a.hstart[sA], a1 = sA + 1 < T ? a.hstart[sA + 1] : M;
                const int b0 = a.hstart[sB], b1 = sB + 1 < T ? a.hstart[sB + 1] : M;
                float lower, upper;
                box_bounds(a.fbox[2 * (size_t)a0], a.fbox[2 * (size_t)a0 + 1], a.fbox[2 * (size_t)b0], a.fbox[2 * (size_t)b0 + 1], lower, upper);
                if (lower < a.r2) {              // else no pair of the two boxes can satisfy d2 < r2 (the bound is exact in fp32, cell_uf.cuh)
                    bool hit = upper < a.r2;     // every pair does
                    for (int i = a0; i < a1 && !hit; ++i) {
                        const float4 p = a.spts[i];
                        for (int j = b0; j < b1; ++j) {
                            const float4 q = a.spts[j];
                            if (dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z) < a.r2) { hit = true; break; }
                        }
                    }
                    if (hit) { lo = min(a0, b0); hi = max(a0, b0); }
                }
            }
        }
        // append the confirmed edges of the warp with one atomic
        const unsigned m = __ballot_sync(kFull, hi >= 0);
        if (m) {
            int base = 0;
            if (lane_id() == __ffs(m) - 1) base = atomicAdd(a.state + FS_ST_EDGES, __popc(m));
            base = __shfl_sync(kFull, base, __ffs(m) - 1);
            if (hi >= 0) {
                const int e = base + __popc(m & lanemask_lt());
                if (e < a.edge_cap) a.edges[e] = (uint32_t)lo | ((uint32_t)hi << 16);
                else atomicOr(a.counts + CNT_FLAGS, FS_FLAG_FALLBACK);
                atomicMin(a.parent + hi, lo);  // result unused: a reduction, no round trip
            }
        }
    }
}

// lock-free union in a shared-memory forest, both walks in step (two independent loads per round); the smaller id stays root
__device__ __forceinline__ void fs_smem_union(int* tp, int u, int v) {
    for (;;) {
        const int pu = tp[u], pv = tp[v];
        if (pu == pv) return;  // same parent: same tree
        const int gu = tp[pu], gv = tp[pv];
        if (gu != pu || gv != pv) {  // not both at a root yet: halve the paths and go on
            if (gu != pu) tp[u] = gu;
            if (gv != pv) tp[v] = gv;
            u = gu;
            v = gv;
            continue;
        }
        const int lo = min(pu, pv), hi = max(pu, pv);
        const int old = atomicCAS(tp + hi, hi, lo);
        if (old == hi) return;
        u = old;  // hi had been hooked meanwhile: join where it went
        v = lo;
    }
}

// ======================================================================================================================
// Kernel 3 (one cluster): flatten, sizes, kept clusters in PCL's order, CSR with ascending indices.
// ======================================================================================================================
__global__ void __launch_bounds__(FS_THREADS, 1) k_fs_tables(const FsArgs* __restrict__ ap) {
    const FsArgs a = *ap;
    cg::cluster_group cl = cg::this_cluster();
    __shared__ int sscan[33];
    __shared__ unsigned long long ssum64[32];
    extern __shared__ __align__(16) int dyn[];  // FS_TABLES_SMEM bytes: the parent array first, the cluster tables (sbuf) afterwards
    int* sbuf = dyn;                            // [2 * FS_MAX_K + 2]
    const int tid = threadIdx.x;
    const int rank = (int)cl.block_rank(), nb = (int)cl.num_blocks();
    const int NT = nb * FS_THREADS;
    const int gtid = rank * FS_THREADS + tid;
    const int lane = lane_id();
    fs_stamp(a, gtid, 6);
    if (a.counts[CNT_FLAGS] != 0) return;
    const int M = a.counts[CNT_M];

    // ---- E2: union-find over the listed cell pairs in SHARED memory, by every CTA for itself -------------------------------
    // The first version joined the cells with finds and CAS in global memory from a wide kernel: every step a dependent L2 round
    // trip, most of them on the few lines that hold the roots of the large components (ncu: 45-60 % of the stall samples in the
    // find loops), 20-30 us for 95 k edges.  A small frame's parent array (one int per kept point, node id = cell start) fits the
    // 192 KB of one CTA, so each CTA of the cluster runs the WHOLE union-find in its own shared memory -- redundantly, but with no
    // merge and ~30-cycle steps -- and flattens its share of the points against it.  (Measured alternative: every CTA joins 1 / nb
    // of the edges and the forests are merged pairwise through distributed shared memory -- 12-17 us per merge round, 65 us.)
    // The forest does not start from singletons: k_fs_edges has hooked every cell under its smallest confirmed neighbour
    // (parent[], one fire-and-forget atomicMin per edge), so after one flattening pass most listed pairs already share a parent
    // and leave the loop below at its first comparison.
    int* tp = dyn;  // [M]
    for (int x = tid; x < M; x += FS_THREADS) tp[x] = a.parent[x];
    __syncthreads();
    for (int x = tid; x < M; x += FS_THREADS) {  // a parent is only ever replaced by an ancestor: concurrent walks stay correct
        int r = tp[x];
        if (r == x) continue;
        while (tp[r] != r) r = tp[r];
        tp[x] = r;
    }
    __syncthreads();
    fs_stamp(a, gtid, 16);
    {
        const int n_edges = a.state[FS_ST_EDGES];
        if (gtid == 0) a.phase_ns[FS_PHASES - 1] = (unsigned long long)n_edges;  // (diagnostics: listed cell pairs)
        const uint32_t* ed = a.edges;
        uint32_t n0 = tid < n_edges ? ed[tid] : 0u, n1 = tid + FS_THREADS < n_edges ? ed[tid + FS_THREADS] : 0u;
        for (int e = tid; e < n_edges; e += FS_THREADS) {
            const uint32_t cur = n0;
            n0 = n1;
            if (e + 2 * FS_THREADS < n_edges) n1 = ed[e + 2 * FS_THREADS];  // two loads in flight while this edge is processed
            fs_smem_union(tp, (int)(cur & 0xffffu), (int)(cur >> 16));
        }
        __syncthreads();
    }
    fs_stamp(a, gtid, 17);

    // ---- F: flatten, component size, smallest original index ---------------------------------------------------------------
    // Sizes and minima are first gathered per CTA in shared memory (behind the parent array, when 3 M ints fit) and flushed with
    // one global atomic per (CTA, root): thousands of points of one large component otherwise queue up on a single L2 address
    // (12 us of the 30 this kernel took on a c1 frame).
    const bool local_acc = 3 * (size_t)M * 4 <= FS_TABLES_SMEM;  // uniform
    int* scnt = dyn + M;
    int* smin = dyn + 2 * M;
    if (local_acc) {
        for (int x = tid; x < M; x += FS_THREADS) {
            scnt[x] = 0;
            smin[x] = 0x7fffffff;
        }
        __syncthreads();
    }
    const int rounds = (M + NT - 1) / NT;
    for (int it = 0; it < rounds; ++it) {
        const int s = it * NT + gtid;
        const bool valid = s < M;
        int r = -1, orig = 0x7fffffff;
        if (valid) {
            r = a.scell[s];
            while (tp[r] != r) r = tp[r];  // (read only: other threads of the CTA are walking too)
            a.root[s] = r;
            orig = __float_as_int(a.spts[s].w);
        }
        const unsigned peers = __match_any_sync(kFull, r);
        const int omin = __reduce_min_sync(peers, orig);
        if (valid && lane == __ffs(peers) - 1) {
            if (local_acc) {
                atomicAdd(scnt + r, __popc(peers));
                atomicMin(smin + r, omin);
            } else {
                atomicAdd(a.csize + r, __popc(peers));
                atomicMin(a.cmin + r, omin);
            }
        }
    }
    if (local_acc) {
        __syncthreads();
        for (int x = tid; x < M; x += FS_THREADS) {
            const int c = scnt[x];
            if (c > 0) {
                atomicAdd(a.csize + x, c);
                atomicMin(a.cmin + x, smin[x]);
            }
        }
    }
    fs_stamp(a, gtid, 18);

    // ---- G: clusters inside [min, max]; labels ---------------------------------------------------------------------------
    cl.sync();
    fs_stamp(a, gtid, 7);
    for (int s = gtid; s < M; s += NT) {
        const int r = a.root[s];
        a.labels[__float_as_int(a.spts[s].w)] = a.cmin[r];
        if (r != s) continue;
        const int sz = a.csize[s];
        if (sz < a.min_size || sz > a.max_size) continue;
        const int k = atomicAdd(a.counts + CNT_K, 1);
        atomicAdd(a.counts + CNT_TOTAL, sz);
        if (k < FS_MAX_K) {
            a.ksize[k] = sz;
            a.kmin[k] = a.cmin[s];
            a.kroot[k] = s;
        }
    }

    // ---- H: rank of every cluster: (size desc, smallest index asc) ----------------------------------------------------------
    cl.sync();
    fs_stamp(a, gtid, 8);
    const int K = ld_cg(a.counts + CNT_K);
    if (K > FS_MAX_K) {  // uniform
        if (gtid == 0) atomicOr(a.counts + CNT_FLAGS, FS_FLAG_FALLBACK);
        return;
    }
    for (int k = tid; k < K; k += FS_THREADS) {
        sbuf[2 * k] = a.ksize[k];
        sbuf[2 * k + 1] = a.kmin[k];
    }
    __syncthreads();
    for (int k = gtid; k < K; k += NT) {
        const int sz = sbuf[2 * k], mi = sbuf[2 * k + 1];
        int before = 0;
        for (int j = 0; j < K; ++j) {
            const int sj = sbuf[2 * j], mj = sbuf[2 * j + 1];
            before += (sj > sz) || (sj == sz && mj < mi);
        }
        a.crank[a.kroot[k]] = before;
        a.ssize[before] = sz;
    }

    // ---- I: CSR offsets, every CTA for itself (shared memory), CTA 0 publishes them -------------------------------------------
    cl.sync();
    fs_stamp(a, gtid, 9);
    int* soff = sbuf;  // [K + 1]
    int total = 0;
    {
        const int per = (K + FS_THREADS - 1) / FS_THREADS;  // <= 4
        const int k0 = tid * per;
        int v[4] = {0, 0, 0, 0};
        int sum = 0;
        unsigned long long sq = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < per && k0 + k < K) {
                v[k] = a.ssize[k0 + k];
                sum += v[k];
                sq += (unsigned long long)v[k] * (unsigned long long)v[k];
            }
        const int excl = block_exclusive_scan(sum, sscan, &total);  // starts with a barrier: sbuf's previous readers are done
        const unsigned long long sq_all = fs_block_sum64(sq, ssum64);
        if (sq_all > a.order_limit) {  // uniform: every CTA computes the same sum
            if (gtid == 0) atomicOr(a.counts + CNT_FLAGS, FS_FLAG_FALLBACK);
            return;
        }
        int run = excl;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < per && k0 + k < K) {
                soff[k0 + k] = run;
                if (rank == 0) a.cl_offsets[k0 + k] = run;
                run += v[k];
            }
        if (tid == 0) {
            soff[K] = total;
            if (rank == 0) a.cl_offsets[K] = total;
        }
        __syncthreads();
    }

    // ---- J: members into their cluster's segment (any order) -------------------------------------------------------------
    fs_stamp(a, gtid, 10);
    for (int s = gtid; s < M; s += NT) {
        const int k = a.crank[a.root[s]];
        if (k < 0) continue;
        const int pos = soff[k] + atomicAdd(a.cursor + k, 1);
        a.idx_tmp[pos] = (uint32_t)__float_as_int(a.spts[s].w);
    }

    // ---- K: ascending original index inside every cluster (rank by counting; bounded by order_limit) ------------------------
    cl.sync();
    fs_stamp(a, gtid, 11);
    for (int t = gtid; t < total; t += NT) {
        int lo = 0, hi = K - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (soff[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int b0 = soff[lo], b1 = soff[lo + 1];
        const uint32_t v = a.idx_tmp[t];
        int before = 0;
        // every lane of a warp reads the same addresses (one L1 wavefront per load): four indices per load
        int u = b0;
        const int body0 = min(b1, (b0 + 3) & ~3), body1 = max(body0, b1 & ~3);
        for (; u < body0; ++u) before += a.idx_tmp[u] < v;
        for (; u < body1; u += 4) {
            const uint4 q = *reinterpret_cast<const uint4*>(a.idx_tmp + u);
            before += (q.x < v) + (q.y < v) + (q.z < v) + (q.w < v);
        }
        for (; u < b1; ++u) before += a.idx_tmp[u] < v;
        a.indices[b0 + before] = v;
    }
    fs_stamp(a, gtid, 12);
}

// ======================================================================================================================
// Kernel 4 (whole GPU): farthest pair of every cluster (getCentroid step 1, MOT.cpp:708-760): `slabs` CTAs per cluster share
// its rows, each stages the cluster's points in shared memory once.
// ======================================================================================================================
constexpr int FS_FP_THREADS = 256;
constexpr int FS_FP_SMEM_POINTS = 2048;
__device__ __forceinline__ int fs_slabs(int ctas, int K) {
    int s = K > 0 ? ctas / K : 1;
    return s < 1 ? 1 : (s > 64 ? 64 : s);
}
__global__ void __launch_bounds__(FS_FP_THREADS) k_fs_farthest(const FsArgs* __restrict__ ap) {
    const FsArgs a = *ap;
    __shared__ float4 sp[FS_FP_SMEM_POINTS];
    __shared__ PairCand sbest[FS_FP_THREADS / 32];
    fs_stamp(a, blockIdx.x * FS_FP_THREADS + threadIdx.x, 13);
    if (a.counts[CNT_FLAGS] != 0 || !a.with_centroids) return;
    const int K = a.counts[CNT_K];
    const float4* cloud = a.do_rs ? a.kept : a.src;
    const int slabs = fs_slabs(a.fp_ctas, K);
    for (int item = blockIdx.x; item < K * slabs; item += gridDim.x) {
        const int c = item / slabs, slab = item % slabs;
        const int s0 = a.cl_offsets[c], n = a.cl_offsets[c + 1] - s0;
        const bool staged = n <= FS_FP_SMEM_POINTS;
        __syncthreads();
        if (staged) {
            for (int t = threadIdx.x; t < n; t += FS_FP_THREADS) sp[t] = cloud[a.indices[s0 + t]];
            __syncthreads();
        }
        PairCand best;
        best.dist = -1.0f; best.i = 0x7fffffff; best.j = 0x7fffffff;
        double seen_s = -1.0;
        const int row_stride = slabs * (FS_FP_THREADS / 32);
        for (int i = slab * (FS_FP_THREADS / 32) + warp_id(); i < n - 1; i += row_stride) {
            const float4 pi = staged ? sp[i] : cloud[a.indices[s0 + i]];
            for (int j = i + 1 + lane_id(); j < n; j += 32) {
                const float4 pj = staged ? sp[j] : cloud[a.indices[s0 + j]];
                pair_scan_step(pi, pj, i, j, best, seen_s);  // (i, j) ascending within a lane: strict > keeps the first
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            PairCand other;
            other.dist = __shfl_xor_sync(kFull, best.dist, o);
            other.i = __shfl_xor_sync(kFull, best.i, o);
            other.j = __shfl_xor_sync(kFull, best.j, o);
            if (cand_better(other, best)) best = other;
        }
        if (lane_id() == 0) sbest[warp_id()] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < FS_FP_THREADS / 32; ++w)
                if (cand_better(sbest[w], best)) best = sbest[w];
            a.cands[item] = best;
        }
    }
}

// ======================================================================================================================
// Kernel 5: statistics (+ line distance + circumcentre, MOT.cpp:761-822) of every cluster, one CTA each.
// ======================================================================================================================
constexpr int FS_FIN_THREADS = 128;
__global__ void __launch_bounds__(FS_FIN_THREADS) k_fs_finish(const FsArgs* __restrict__ ap) {
    const FsArgs a = *ap;
    __shared__ double ssum[FS_FIN_THREADS / 32][3];
    __shared__ float sminmax[FS_FIN_THREADS / 32][6];
    __shared__ float sdist[FS_FIN_THREADS / 32];
    __shared__ int sk[FS_FIN_THREADS / 32];
    __shared__ PairCand s2[2];
    fs_stamp(a, blockIdx.x * FS_FIN_THREADS + threadIdx.x, 14);
    if (blockIdx.x == 0 && threadIdx.x < CNT_N) {  // the counters are final (this kernel changes none of them)
        a.host_counts[threadIdx.x] = a.counts[threadIdx.x];
        __threadfence_system();
    }
    if (a.counts[CNT_FLAGS] != 0) return;
    const int K = a.counts[CNT_K];
    const float4* cloud = a.do_rs ? a.kept : a.src;
    const int lane = lane_id(), w = warp_id();
    const int slabs = fs_slabs(a.fp_ctas, K);
    for (int c = blockIdx.x; c < K; c += gridDim.x) {
        const int s0 = a.cl_offsets[c], n = a.cl_offsets[c + 1] - s0;
        CcLine L;
        if (a.with_centroids) {  // uniform
            const PairCand best = reduce_cands(a.cands + (size_t)c * slabs, slabs, s2);
            const bool have = best.dist >= 0.0f;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            cc_line_from_pair(have ? cloud[a.indices[s0 + best.i]] : z, have ? cloud[a.indices[s0 + best.j]] : z, have, L);
        }
        double sx = 0, sy = 0, sz = 0;
        float n0 = INFINITY, n1 = INFINITY, n2 = INFINITY, x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY;
        float bd = -1.0f;
        int bk = 0x7fffffff;
        for (int k = threadIdx.x; k < n; k += FS_FIN_THREADS) {
            const float4 p = cloud[a.indices[s0 + k]];
            sx += (double)p.x; sy += (double)p.y; sz += (double)p.z;
            n0 = fminf(n0, p.x); n1 = fminf(n1, p.y); n2 = fminf(n2, p.z);
            x0 = fmaxf(x0, p.x); x1 = fmaxf(x1, p.y); x2 = fmaxf(x2, p.z);
            if (a.with_centroids) {
                bool skip;
                const float d = cc_line_dist(L, p, skip);
                if (d > bd && !skip) { bd = d; bk = k; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(kFull, sx, o); sy += __shfl_xor_sync(kFull, sy, o); sz += __shfl_xor_sync(kFull, sz, o);
            n0 = fminf(n0, __shfl_xor_sync(kFull, n0, o)); n1 = fminf(n1, __shfl_xor_sync(kFull, n1, o)); n2 = fminf(n2, __shfl_xor_sync(kFull, n2, o));
            x0 = fmaxf(x0, __shfl_xor_sync(kFull, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(kFull, x1, o)); x2 = fmaxf(x2, __shfl_xor_sync(kFull, x2, o));
            const float od = __shfl_xor_sync(kFull, bd, o);
            const int ok = __shfl_xor_sync(kFull, bk, o);
            if (od > bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
        }
        __syncthreads();
        if (lane == 0) {
            ssum[w][0] = sx; ssum[w][1] = sy; ssum[w][2] = sz;
            sminmax[w][0] = n0; sminmax[w][1] = n1; sminmax[w][2] = n2; sminmax[w][3] = x0; sminmax[w][4] = x1; sminmax[w][5] = x2;
            sdist[w] = bd; sk[w] = bk;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int v = 1; v < FS_FIN_THREADS / 32; ++v) {
                sx += ssum[v][0]; sy += ssum[v][1]; sz += ssum[v][2];
                n0 = fminf(n0, sminmax[v][0]); n1 = fminf(n1, sminmax[v][1]); n2 = fminf(n2, sminmax[v][2]);
                x0 = fmaxf(x0, sminmax[v][3]); x1 = fmaxf(x1, sminmax[v][4]); x2 = fmaxf(x2, sminmax[v][5]);
                if (sdist[v] > bd || (sdist[v] == bd && sk[v] < bk)) { bd = sdist[v]; bk = sk[v]; }
            }
            ClusterStat st;
            st.count = n;
            st.mean[0] = (float)(sx / (double)n); st.mean[1] = (float)(sy / (double)n); st.mean[2] = (float)(sz / (double)n);
            st.bmin[0] = n0; st.bmin[1] = n1; st.bmin[2] = n2;
            st.bmax[0] = x0; st.bmax[1] = x1; st.bmax[2] = x2;
            a.stats[c] = st;
            if (a.with_centroids) {
                double Pk[3] = {0, 0, 0};
                if (bk != 0x7fffffff && bd >= 0.0f) {
                    const float4 p = cloud[a.indices[s0 + bk]];
                    Pk[0] = p.x; Pk[1] = p.y; Pk[2] = p.z;
                }
                a.centroids[c] = cc_finish(L, Pk, a.stamp);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.phase_ns[15] = t;
    }
}

}  // namespace mot
