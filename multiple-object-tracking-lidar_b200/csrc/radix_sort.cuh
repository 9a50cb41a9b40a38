// radix_sort.cuh -- stable LSD radix sort of (key, u32 value) pairs, hand written for sm_100a.
//
// Used three times per frame: (1) points by voxel key (grid build), (2) clusters by
// (size desc, min index asc), (3) the stable partition of point indices by cluster rank (CSR emission, K6).
//
// Per pass (digit width 1..10 bits, chosen by the host so that passes = ceil(key_bits / 10)):
//   k_rs_hist     every block histograms the digit over its contiguous chunk          (reads keys)
//   k_rs_scan     one warp per digit: exclusive prefix of that digit over the blocks  (R x G counters)
//   k_rs_scatter  tile-wise stable ranking with warp match_any + per-warp counters, tile re-ordered in
//                 shared memory so that global writes are coalesced runs              (reads + writes pairs)
// No inter-block spin waits anywhere: a pass is three stream-ordered launches.
#pragma once
#include "common.cuh"
#include "grid_uf.cuh"

namespace mot {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef RS_ITEMS_V
#define RS_ITEMS_V 8
#endif
constexpr int RS_ITEMS = RS_ITEMS_V;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // pairs per tile (2048 at 8 items per thread)
constexpr int RS_MAX_BITS = 10;
#ifndef RS_MAX_GRID_V
#define RS_MAX_GRID_V 1184
#endif
// 8 chunks per SM: more blocks than can be resident, started in index order, so the co-resident blocks fill neighbouring
// ranges of every digit's output region (their partial sectors merge in L2); 592 measured 20 % slower in the scatter,
// 2368 costs more in the counter matrix than it gains
constexpr int RS_MAX_GRID = RS_MAX_GRID_V;

template <typename KT>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const KT* __restrict__ keys, int n, int chunk, int shift, int bits,
                                                         unsigned* __restrict__ hist /* [R][G] */) {
    extern __shared__ unsigned sh_hist[];
    const int R = 1 << bits;
    const unsigned mask = R - 1;
    for (int d = threadIdx.x; d < R; d += RS_THREADS) sh_hist[d] = 0;
    __syncthreads();
    const int begin = blockIdx.x * chunk;
    const int end = min(n, begin + chunk);
    const int lane = lane_id();
    // RS_ITEMS coalesced loads in flight per thread, then run-aggregated shared-memory atomics (match.any per key measured
    // 1.6x slower here: 49 -> 31 us per pass on 16 M keys)
    for (int base = begin; base < end; base += RS_TILE) {
        KT k[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const int i = base + j * RS_THREADS + threadIdx.x;
            k[j] = i < end ? keys[i] : (KT)0;
        }
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const int i = base + j * RS_THREADS + threadIdx.x;
            const bool valid = i < end;
            const unsigned d = valid ? (unsigned)((k[j] >> shift) & mask) : 0xffffffffu;
            // runs of equal digits in neighbouring lanes (the common shape of conflicts: sorted or nearly constant digits)
            // collapse to one shared-memory atomic per run; scattered digits cost one atomic each
            const unsigned dp = __shfl_up_sync(kFull, d, 1);
            const unsigned heads = __ballot_sync(kFull, lane == 0 || d != dp);
            if (heads >> lane & 1u) {
                const unsigned later = heads & ~((2u << lane) - 1u);  // heads after this lane
                const int run = (later ? __ffs(later) - 1 : 32) - lane;
                if (valid) atomicAdd(&sh_hist[d], (unsigned)run);
            }
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < R; d += RS_THREADS) hist[(size_t)d * gridDim.x + blockIdx.x] = sh_hist[d];
}

// One warp per digit: prefix[d][b] = sum_{b' < b} hist[d][b'], tot[d] = sum_b hist[d][b].
__global__ void __launch_bounds__(256) k_rs_scan(const unsigned* __restrict__ hist, unsigned* __restrict__ prefix,
                                                  unsigned* __restrict__ tot, int R, int G) {
    const int d = blockIdx.x * 8 + warp_id();
    if (d >= R) return;
    const int lane = lane_id();
    const unsigned* row = hist + (size_t)d * G;
    unsigned* prow = prefix + (size_t)d * G;
    // 32 consecutive counters per step (coalesced 128-byte loads / stores), one shuffle scan per step, the running total carried
    // in a register; two steps' loads are in flight together (round 1 gave every lane a contiguous run of counters: strided
    // 4-byte accesses and 2 x ceil(G / 32) dependent loads per lane -- 21 us per pass where this takes a few)
    unsigned run = 0;
    for (int c = 0; c < G; c += 64) {
        const int i0 = c + lane, i1 = c + 32 + lane;
        const int v0 = i0 < G ? (int)row[i0] : 0, v1 = i1 < G ? (int)row[i1] : 0;
        const int s0 = warp_inclusive_scan(v0), s1 = warp_inclusive_scan(v1);
        const unsigned t0 = (unsigned)__shfl_sync(kFull, s0, 31);
        if (i0 < G) prow[i0] = run + (unsigned)(s0 - v0);
        if (i1 < G) prow[i1] = run + t0 + (unsigned)(s1 - v1);
        run += t0 + (unsigned)__shfl_sync(kFull, s1, 31);
    }
    if (lane == 0) tot[d] = run;
}

#ifndef RS_BIG_THREADS_V
#define RS_BIG_THREADS_V 512
#endif
constexpr int RS_BIG_THREADS = RS_BIG_THREADS_V;   // CTA width of the scatter on large inputs
constexpr int RS_BIG_TILE = 8192;                  // pairs per tile there: 4x longer digit runs per tile (fewer partial sectors at L2)
constexpr int RS_ITEMS_BIG = RS_BIG_TILE / RS_BIG_THREADS;
#ifndef RS_SCATTER_CTAS_BIG
#define RS_SCATTER_CTAS_BIG 2
#endif
// shared memory of the scatter: tile_off[R], gbase[R], scan scratch, then ONE region that first holds the per-warp digit counters
// and, once every pair knows its slot, the re-ordered tile (the counters are dead by then): 72 KB instead of 106 KB at 8192 pairs,
// so three CTAs fit an SM instead of two
constexpr size_t rs_scatter_smem_bytes(int bits, size_t key_bytes, int items = RS_ITEMS, int threads = RS_THREADS) {
    const size_t cnt = (size_t)(threads / 32) * ((size_t)1 << bits) * 4, tile = (size_t)threads * items * (4 + key_bytes);
    return 2 * ((size_t)1 << bits) * 4 + 36 * 4 + (cnt > tile ? cnt : tile);
}

template <typename KT, bool IOTA, int ITEMS = RS_ITEMS, int THREADS = RS_THREADS>
__global__ void __launch_bounds__(THREADS, THREADS * ITEMS >= RS_BIG_TILE ? RS_SCATTER_CTAS_BIG : 4) k_rs_scatter(const KT* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                            KT* __restrict__ kout, uint32_t* __restrict__ vout, int n, int chunk,
                                                            int shift, int bits, const unsigned* __restrict__ prefix,
                                                            const unsigned* __restrict__ tot) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    const int R = 1 << bits;
    const unsigned mask = R - 1;
    unsigned* tile_off = reinterpret_cast<unsigned*>(rs_smem);  // [R]
    unsigned* gbase = tile_off + R;                             // [R]
    int* scan_tmp = reinterpret_cast<int*>(gbase + R);          // [36]
    unsigned* cnt = reinterpret_cast<unsigned*>(scan_tmp + 36);  // [WARPS][R]   -- the same bytes as --
    uint32_t* st_vals = reinterpret_cast<uint32_t*>(scan_tmp + 36);  // [TILE]
    KT* st_keys = reinterpret_cast<KT*>(st_vals + TILE);             // [TILE]

    const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const int G = gridDim.x, b = blockIdx.x;
    const int per = (R + THREADS - 1) / THREADS;  // digits owned by a thread (<= 4)
    const int d0 = tid * per;

    // gbase[d] = (exclusive scan of tot over digits)[d] + prefix[d][b]
    {
        int local = 0;
        for (int k = 0; k < per; ++k) {
            const int d = d0 + k;
            if (d < R) local += (int)tot[d];
        }
        int total;
        int run = block_exclusive_scan(local, scan_tmp, &total);
        for (int k = 0; k < per; ++k) {
            const int d = d0 + k;
            if (d < R) {
                gbase[d] = (unsigned)run + prefix[(size_t)d * G + b];
                run += (int)tot[d];
            }
        }
    }
    __syncthreads();

    const int begin = b * chunk;
    const int end = min(n, begin + chunk);
    for (int tile_begin = begin; tile_begin < end; tile_begin += TILE) {
        {   // clear the counters (the previous tile's write-out has been fenced by the barrier that ends the loop body)
            uint4* c4 = reinterpret_cast<uint4*>(cnt);
            for (int i = tid; i < WARPS * R / 4; i += THREADS) c4[i] = make_uint4(0u, 0u, 0u, 0u);
            if (R < 4)
                for (int i = tid; i < WARPS * R; i += THREADS) cnt[i] = 0;
        }
        __syncthreads();

        KT key[ITEMS];
        unsigned rank[ITEMS / 2];  // two 16-bit slots per register (a slot is < TILE <= 65536)
        static_assert(ITEMS % 8 == 0 && THREADS * ITEMS <= 65536, "packed ranks, batches of eight");
        const int seg = tile_begin + w * (32 * ITEMS);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int idx = seg + i * 32 + lane;
            key[i] = idx < end ? kin[idx] : ~(KT)0;
        }
        unsigned* wcnt = cnt + w * R;
#ifndef RS_RANK_LDST
        // Rank inside the warp: match.any groups the lanes with equal digits, the group's first lane bumps the warp's counter with ONE
        // shared-memory atomic.  Eight items are in flight: the atomics of a batch are issued back to back (the counter update of
        // item i+1 does not wait for the value item i got back, as the load / add / store version did -- the loop was one long
        // dependency chain through shared memory), their results are collected afterwards.  Atomics of one warp on one address
        // are applied in program order (__syncwarp between them), so the ranks stay stable.
#pragma unroll
        for (int i0 = 0; i0 < ITEMS; i0 += 8) {
            unsigned peers[8], old[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int idx = seg + (i0 + j) * 32 + lane;
                const unsigned d = idx < end ? (unsigned)((key[i0 + j] >> shift) & mask) : mask;  // padding sorts last
                peers[j] = __match_any_sync(kFull, d);
                old[j] = 0;
                if (lane == __ffs(peers[j]) - 1) old[j] = atomicAdd(wcnt + d, (unsigned)__popc(peers[j]));
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = i0 + j;
                const unsigned r = __shfl_sync(kFull, old[j], __ffs(peers[j]) - 1) + (unsigned)__popc(peers[j] & lanemask_lt());
                if (i & 1) rank[i / 2] |= r << 16;
                else rank[i / 2] = r;
            }
        }
#else
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int idx = seg + i * 32 + lane;
            const unsigned d = idx < end ? (unsigned)((key[i] >> shift) & mask) : mask;  // padding sorts last
            const unsigned peers = __match_any_sync(kFull, d);
            const int leader = __ffs(peers) - 1;
            unsigned old = 0;
            if (lane == leader) {
                old = wcnt[d];
                wcnt[d] = old + (unsigned)__popc(peers);
            }
            old = __shfl_sync(kFull, old, leader);
            const unsigned r = old + (unsigned)__popc(peers & lanemask_lt());
            if (i & 1) rank[i / 2] |= r << 16;
            else rank[i / 2] = r;
            __syncwarp();
        }
#endif
        __syncthreads();

        // per digit: exclusive prefix over the warps; tile_off[d] temporarily holds the tile's digit count
        for (int d = tid; d < R; d += THREADS) {
            unsigned run = 0;
#pragma unroll
            for (int ww = 0; ww < WARPS; ++ww) {
                const unsigned t = cnt[ww * R + d];
                cnt[ww * R + d] = run;
                run += t;
            }
            tile_off[d] = run;
        }
        __syncthreads();
        unsigned mycount[4] = {0, 0, 0, 0};
        {
            int local = 0;
            for (int k = 0; k < per; ++k) {
                const int d = d0 + k;
                if (d < R) {
                    mycount[k] = tile_off[d];
                    local += (int)mycount[k];
                }
            }
            int total;
            int run = block_exclusive_scan(local, scan_tmp, &total);
            for (int k = 0; k < per; ++k) {
                const int d = d0 + k;
                if (d < R) {
                    tile_off[d] = (unsigned)run;
                    run += (int)mycount[k];
                }
            }
        }
        __syncthreads();

        // slot of every pair inside the re-ordered tile; after the barrier the counters are dead and their bytes take the tile
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int idx = seg + i * 32 + lane;
            const unsigned d = idx < end ? (unsigned)((key[i] >> shift) & mask) : mask;
            rank[i / 2] += (tile_off[d] + wcnt[d]) << (16 * (i & 1));  // (no carry out of a half: the sum is a slot of this tile)
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int idx = seg + i * 32 + lane;
            const unsigned slot = (rank[i / 2] >> (16 * (i & 1))) & 0xffffu;
            st_keys[slot] = key[i];
            st_vals[slot] = IOTA ? (uint32_t)idx : (idx < end ? vin[idx] : 0u);  // (values are only read now: they never sat in registers beside keys and ranks)
        }
        __syncthreads();

        const int tile_n = min(TILE, end - tile_begin);
        for (int j = tid; j < tile_n; j += THREADS) {
            const KT k = st_keys[j];
            const unsigned d = (unsigned)((k >> shift) & mask);
            const unsigned g = gbase[d] + ((unsigned)j - tile_off[d]);
            kout[g] = k;
            vout[g] = st_vals[j];
        }
        __syncthreads();
        for (int k = 0; k < per; ++k) {
            const int d = d0 + k;
            if (d < R) gbase[d] += mycount[k];
        }
        __syncthreads();
    }
}

// ---- onesweep: one kernel per pass, decoupled look-back instead of the per-block histogram matrix ----------------
//
// k_os_hist reads the keys once and accumulates the GLOBAL digit histogram of every pass.  A pass is then a single
// launch of k_os_pass: a CTA takes the next tile (ticket from an atomic counter, so every predecessor tile is already
// running or done), ranks its 4096 pairs (match.any + per-warp counters, stable), publishes the tile's digit counts,
// obtains its exclusive prefix per digit by walking back over the predecessors' status words (aggregate / inclusive
// prefix packed with a 2-bit flag in one 32-bit word, so no fences are needed), and scatters through shared memory.
// Keys move HBM -> SM -> HBM once per pass; nothing else is re-read.
constexpr int OS_THREADS = 256;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_ITEMS = 8;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;  // 2048 pairs per tile
constexpr int OS_MAX_PASSES = 7;
constexpr unsigned OS_FLAG_AGG = 1u << 30, OS_FLAG_INCL = 2u << 30, OS_VALUE_MASK = (1u << 30) - 1u;
constexpr int OS_SPIN_LIMIT = 1 << 24;  // a stuck look-back raises an error flag instead of hanging the GPU

struct OsPlan {
    int passes;
    int shift[OS_MAX_PASSES];
    int bits[OS_MAX_PASSES];
};
inline OsPlan os_plan(int key_bits) {
    OsPlan p;
    if (key_bits < 1) key_bits = 1;
    p.passes = (key_bits + RS_MAX_BITS - 1) / RS_MAX_BITS;
    const int base = key_bits / p.passes, rem = key_bits % p.passes;
    int sh = 0;
    for (int i = 0; i < p.passes; ++i) {
        p.bits[i] = base + (i < rem ? 1 : 0);
        p.shift[i] = sh;
        sh += p.bits[i];
    }
    return p;
}

// ghist[pass][1 << RS_MAX_BITS]; one launch covers up to 4 passes (pass p0 .. p0+NP-1), shifts/widths in registers
struct OsHistArgs { int shift[4]; int bits[4]; };
template <typename KT, int NP>
__global__ void __launch_bounds__(RS_THREADS) k_os_hist(const KT* __restrict__ keys, int n, int chunk, OsHistArgs a, unsigned* __restrict__ ghist) {
    extern __shared__ unsigned sh_hist[];  // [NP][1 << RS_MAX_BITS]
    constexpr int RMAX = 1 << RS_MAX_BITS;
    for (int d = threadIdx.x; d < NP * RMAX; d += RS_THREADS) sh_hist[d] = 0;
    __syncthreads();
    const int begin = blockIdx.x * chunk;
    const int end = min(n, begin + chunk);
    const int lane = lane_id();
    for (int base = begin; base < end; base += RS_TILE) {
        KT k[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const int i = base + j * RS_THREADS + threadIdx.x;
            k[j] = i < end ? keys[i] : (KT)0;
        }
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const int i = base + j * RS_THREADS + threadIdx.x;
            const bool valid = i < end;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const unsigned d = valid ? (unsigned)((k[j] >> a.shift[p]) & (KT)((1u << a.bits[p]) - 1u)) : 0xffffffffu;
                const unsigned peers = __match_any_sync(kFull, d);
                if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh_hist[p * RMAX + d], (unsigned)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < NP * RMAX; d += RS_THREADS) {
        const unsigned v = sh_hist[d];
        if (v) atomicAdd(&ghist[d], v);
    }
}

template <typename KT>
inline void os_launch_hist(cudaStream_t st, const KT* keys, int n, const OsPlan& plan, unsigned* ghist, Prof& prof, int kid) {
    constexpr int RMAX = 1 << RS_MAX_BITS;
    const Chunking hk = make_chunking(n, RS_TILE, RS_MAX_GRID);
    for (int p0 = 0; p0 < plan.passes; p0 += 4) {
        const int np = plan.passes - p0 < 4 ? plan.passes - p0 : 4;
        OsHistArgs a{};
        for (int p = 0; p < np; ++p) { a.shift[p] = plan.shift[p0 + p]; a.bits[p] = plan.bits[p0 + p]; }
        const size_t smem = (size_t)np * RMAX * sizeof(unsigned);
        unsigned* gh = ghist + (size_t)p0 * RMAX;
        prof.begin(kid);
        if (np == 1) k_os_hist<KT, 1><<<hk.grid, RS_THREADS, smem, st>>>(keys, n, hk.chunk, a, gh);
        else if (np == 2) k_os_hist<KT, 2><<<hk.grid, RS_THREADS, smem, st>>>(keys, n, hk.chunk, a, gh);
        else if (np == 3) k_os_hist<KT, 3><<<hk.grid, RS_THREADS, smem, st>>>(keys, n, hk.chunk, a, gh);
        else k_os_hist<KT, 4><<<hk.grid, RS_THREADS, smem, st>>>(keys, n, hk.chunk, a, gh);
        prof.end();
    }
}

constexpr size_t os_pass_smem_bytes(int bits, size_t key_bytes) {
    return (size_t)(OS_WARPS + 2) * ((size_t)1 << bits) * 4 + 40 * 4 + (size_t)OS_TILE * 4 + (size_t)OS_TILE * key_bytes;
}

template <typename KT, bool IOTA>
__global__ void __launch_bounds__(OS_THREADS) k_os_pass(const KT* __restrict__ kin, const uint32_t* __restrict__ vin, KT* __restrict__ kout,
                                                         uint32_t* __restrict__ vout, int n, int shift, int bits,
                                                         const unsigned* __restrict__ ghist, unsigned* status, unsigned* tile_counter,
                                                         int* err_flag) {
    extern __shared__ __align__(16) unsigned char os_smem[];
    const int R = 1 << bits;
    const unsigned mask = R - 1;
    unsigned* cnt = reinterpret_cast<unsigned*>(os_smem);  // [OS_WARPS][R]
    unsigned* tile_off = cnt + OS_WARPS * R;               // [R]
    unsigned* gbase = tile_off + R;                        // [R]
    int* scan_tmp = reinterpret_cast<int*>(gbase + R);     // [40]; [36] holds the tile ticket
    uint32_t* st_vals = reinterpret_cast<uint32_t*>(scan_tmp + 40);
    KT* st_keys = reinterpret_cast<KT*>(st_vals + OS_TILE);

    const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const int per = (R + OS_THREADS - 1) / OS_THREADS;  // digits owned by a thread (<= 4)
    const int d0 = tid * per;

    if (tid == 0) scan_tmp[36] = (int)atomicAdd(tile_counter, 1u);
    for (int i = tid; i < OS_WARPS * R; i += OS_THREADS) cnt[i] = 0;
    // first key of every digit in the output = exclusive scan of the global histogram
    {
        int local = 0;
        for (int k = 0; k < per; ++k) {
            const int d = d0 + k;
            if (d < R) local += (int)ghist[d];
        }
        int total;
        int run = block_exclusive_scan(local, scan_tmp, &total);
        for (int k = 0; k < per; ++k) {
            const int d = d0 + k;
            if (d < R) {
                gbase[d] = (unsigned)run;
                run += (int)ghist[d];
            }
        }
    }
    __syncthreads();
    const int tile = scan_tmp[36];
    const int tile_begin = tile * OS_TILE;
    if (tile_begin >= n) return;
    const int end = min(n, tile_begin + OS_TILE);

    KT key[OS_ITEMS];
    uint32_t val[OS_ITEMS];
    unsigned short rank[OS_ITEMS];
    const int seg = tile_begin + w * (32 * OS_ITEMS);
#pragma unroll
    for (int i = 0; i < OS_ITEMS; ++i) {
        const int idx = seg + i * 32 + lane;
        const bool valid = idx < end;
        key[i] = valid ? kin[idx] : ~(KT)0;
        if (IOTA) val[i] = (uint32_t)idx;
        else val[i] = valid ? vin[idx] : 0u;
    }
    unsigned* wcnt = cnt + w * R;
#pragma unroll
    for (int i = 0; i < OS_ITEMS; ++i) {
        const int idx = seg + i * 32 + lane;
        const unsigned d = idx < end ? (unsigned)((key[i] >> shift) & mask) : mask;  // padding sorts last
        const unsigned peers = __match_any_sync(kFull, d);
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (lane == leader) {
            old = wcnt[d];
            wcnt[d] = old + (unsigned)__popc(peers);
        }
        old = __shfl_sync(kFull, old, leader);
        rank[i] = (unsigned short)(old + (unsigned)__popc(peers & lanemask_lt()));
        __syncwarp();
    }
    __syncthreads();

    // per digit: exclusive prefix over the warps, tile count, publish, look back
    const int n_pad = tile_begin + OS_TILE - end;  // padding items were counted under digit `mask`
    unsigned mycount[4] = {0, 0, 0, 0};
    for (int k = 0; k < per; ++k) {  // publish every digit's tile count first ...
        const int d = d0 + k;
        if (d >= R) break;
        unsigned run = 0;
#pragma unroll
        for (int ww = 0; ww < OS_WARPS; ++ww) {
            const unsigned t = cnt[ww * R + d];
            cnt[ww * R + d] = run;
            run += t;
        }
        mycount[k] = run;
        const unsigned real = run - ((unsigned)d == mask ? (unsigned)n_pad : 0u);
        __stcg(status + (size_t)tile * R + d, (tile == 0 ? OS_FLAG_INCL : OS_FLAG_AGG) | real);
    }
    if (tile > 0) {
        for (int k = 0; k < per; ++k) {  // ... then walk back over the predecessors
            const int d = d0 + k;
            if (d >= R) break;
            const unsigned real = mycount[k] - ((unsigned)d == mask ? (unsigned)n_pad : 0u);
            unsigned excl = 0;
            int t = tile - 1, spins = 0;
            for (;;) {
                const unsigned sv = __ldcg(status + (size_t)t * R + d);
                const unsigned f = sv >> 30;
                if (f == 0) {
                    if (++spins > OS_SPIN_LIMIT) { atomicOr(err_flag, 2); break; }
                    continue;
                }
                excl += sv & OS_VALUE_MASK;
                if (f == 2 || t == 0) break;
                --t;
            }
            __stcg(status + (size_t)tile * R + d, OS_FLAG_INCL | (excl + real));
            gbase[d] += excl;
        }
    }
    // exclusive scan of the tile's digit counts -> position of each digit run inside the staged tile
    {
        int local = 0;
        for (int k = 0; k < per; ++k) local += (int)mycount[k];
        int total;
        int run = block_exclusive_scan(local, scan_tmp, &total);
        for (int k = 0; k < per; ++k) {
            const int d = d0 + k;
            if (d < R) {
                tile_off[d] = (unsigned)run;
                run += (int)mycount[k];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < OS_ITEMS; ++i) {
        const int idx = seg + i * 32 + lane;
        const unsigned d = idx < end ? (unsigned)((key[i] >> shift) & mask) : mask;
        const unsigned pos = tile_off[d] + wcnt[d] + rank[i];
        st_keys[pos] = key[i];
        st_vals[pos] = val[i];
    }
    __syncthreads();
    const int tile_n = end - tile_begin;
    for (int j = tid; j < tile_n; j += OS_THREADS) {
        const KT k = st_keys[j];
        const unsigned d = (unsigned)((k >> shift) & mask);
        const unsigned gpos = gbase[d] + ((unsigned)j - tile_off[d]);
        kout[gpos] = k;
        vout[gpos] = st_vals[j];
    }
}

struct RadixWorkspace {
    unsigned* hist = nullptr;    // [R_max][G_max]
    unsigned* prefix = nullptr;  // [R_max][G_max]
    unsigned* tot = nullptr;     // [R_max]
    // onesweep
    unsigned* ghist = nullptr;   // [OS_MAX_PASSES][R_max] + OS_MAX_PASSES tile counters (zeroed together)
    unsigned* status = nullptr;  // [passes][tiles][R]
    size_t status_words = 0;
    int* err_flag = nullptr;     // device int (bit 1: look-back spin limit hit)
    int mode = 0;                // 0 = three kernels per pass (default: faster at 1M-8M pairs on B200), 1 = onesweep
    int digit_bits = RS_MAX_BITS;  // widest digit of the three-kernel path (MOT_SORT_BITS, 4..10)
    int big_tile_from = 1 << 22;   // inputs of at least this many pairs use 8192-pair scatter tiles (MOT_SORT_BIGTILE; 0 = never)
    bool first_hist_done = false;  // the caller already produced pass 0's histogram (k_cell_keys_hist)
};
constexpr size_t rs_workspace_counters() { return (size_t)(1 << RS_MAX_BITS) * RS_MAX_GRID; }

template <typename KT>
inline cudaError_t rs_configure() {
    cudaError_t e;
    {
        const int osm = (int)os_pass_smem_bytes(RS_MAX_BITS, sizeof(KT));
        e = cudaFuncSetAttribute(k_os_pass<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, osm);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_os_pass<KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, osm);
        if (e != cudaSuccess) return e;
    }
    const int smem = (int)rs_scatter_smem_bytes(RS_MAX_BITS, sizeof(KT));
    e = cudaFuncSetAttribute(k_rs_scatter<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_rs_scatter<KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int smem_big = (int)rs_scatter_smem_bytes(RS_MAX_BITS, sizeof(KT), RS_ITEMS_BIG, RS_BIG_THREADS);
    e = cudaFuncSetAttribute(k_rs_scatter<KT, true, RS_ITEMS_BIG, RS_BIG_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_big);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_rs_scatter<KT, false, RS_ITEMS_BIG, RS_BIG_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_big);
}

// pass / chunk plan of the three-kernel path (shared with k_cell_keys_hist, which produces pass 0's histogram itself)
struct RsPlan {
    int passes, base, rem;
    bool big;
    Chunking ck;
    int bits0;  // digit width of pass 0
};
inline RsPlan rs_plan(int n, int key_bits, const RadixWorkspace& ws) {
    RsPlan p;
    if (key_bits < 1) key_bits = 1;
    p.passes = (key_bits + ws.digit_bits - 1) / ws.digit_bits;
    p.base = key_bits / p.passes;
    p.rem = key_bits % p.passes;
    p.big = ws.big_tile_from > 0 && n >= ws.big_tile_from;
    p.ck = make_chunking(n, p.big ? RS_BIG_TILE : RS_TILE, RS_MAX_GRID);
    p.bits0 = p.base + (0 < p.rem ? 1 : 0);
    return p;
}

// K1 + the first histogram of K2 in one pass: voxel key per point (fp64 cell coordinates, as k_cell_keys) written once and
// counted for pass 0's digit while it is still in a register -- the separate k_rs_hist launch re-read all keys for that.
// Same chunking as the sort (block b owns [b*chunk, (b+1)*chunk)), same histogram layout hist[d][b].
template <typename KT>
__global__ void __launch_bounds__(RS_THREADS) k_cell_keys_hist(const float4* __restrict__ pts, int m, int chunk, GridCodec g,
                                                                const int* __restrict__ frame_offsets, KT* __restrict__ keys, int bits,
                                                                unsigned* __restrict__ hist /* [R][G] */, int* __restrict__ flags) {
    extern __shared__ unsigned sh_hist[];
    const int R = 1 << bits;
    const unsigned mask = R - 1;
    for (int d = threadIdx.x; d < R; d += RS_THREADS) sh_hist[d] = 0;
    __syncthreads();
    const int begin = blockIdx.x * chunk;
    const int end = min(m, begin + chunk);
    const int lane = lane_id();
    for (int base = begin; base < end; base += RS_THREADS * 4) {
        float4 p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = base + j * RS_THREADS + threadIdx.x;
            if (i < end) p[j] = ld_stream(pts + i);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = base + j * RS_THREADS + threadIdx.x;
            const bool valid = i < end;
            unsigned d = 0xffffffffu;
            if (valid) {
                int ix, iy, iz;
                cell_coords_checked(g, p[j], ix, iy, iz, flags);
                const int frame = g.n_frames > 1 ? frame_of(frame_offsets, g.n_frames, i) : 0;
                const KT key = key_compose<KT>(g, frame, ix, iy, iz);
                keys[i] = key;
                d = (unsigned)(key & (KT)mask);
            }
            const unsigned dp = __shfl_up_sync(kFull, d, 1);
            const unsigned heads = __ballot_sync(kFull, lane == 0 || d != dp);
            if (heads >> lane & 1u) {
                const unsigned later = heads & ~((2u << lane) - 1u);
                const int run = (later ? __ffs(later) - 1 : 32) - lane;
                if (valid) atomicAdd(&sh_hist[d], (unsigned)run);
            }
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < R; d += RS_THREADS) hist[(size_t)d * gridDim.x + blockIdx.x] = sh_hist[d];
}

// Sorts n pairs by the low `key_bits` bits of the key.  Input in (k[0], v[0]) -- v[0] is ignored and treated
// as 0..n-1 when iota_first.  Returns the index (0/1) of the buffer pair that holds the result.
// Every launch goes through prof (launch counter + optional event timing); kernel ids kid_base + {0,1,2}.
template <typename KT>
inline int radix_sort_pairs(cudaStream_t st, KT* k[2], uint32_t* v[2], int n, int key_bits, bool iota_first,
                            const RadixWorkspace& ws, Prof& prof, int kid_base) {
    if (key_bits < 1) key_bits = 1;
    if (ws.mode == 1 && n > 0) {
        constexpr int RMAX = 1 << RS_MAX_BITS;
        const OsPlan plan = os_plan(key_bits);
        const int tiles = (n + OS_TILE - 1) / OS_TILE;
        size_t need = 0;
        for (int p = 0; p < plan.passes; ++p) need += (size_t)tiles << plan.bits[p];
        if (plan.passes <= OS_MAX_PASSES && need <= ws.status_words) {
            unsigned* counters = ws.ghist + (size_t)OS_MAX_PASSES * RMAX;
            cudaMemsetAsync(ws.ghist, 0, ((size_t)OS_MAX_PASSES * RMAX + OS_MAX_PASSES) * sizeof(unsigned), st);
            cudaMemsetAsync(ws.status, 0, need * sizeof(unsigned), st);
            os_launch_hist<KT>(st, k[0], n, plan, ws.ghist, prof, kid_base);
            int cur = 0;
            size_t soff = 0;
            for (int p = 0; p < plan.passes; ++p) {
                const size_t smem = os_pass_smem_bytes(plan.bits[p], sizeof(KT));
                prof.begin(kid_base + 2);
                if (p == 0 && iota_first)
                    k_os_pass<KT, true><<<tiles, OS_THREADS, smem, st>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, plan.shift[p], plan.bits[p],
                                                                         ws.ghist + (size_t)p * RMAX, ws.status + soff, counters + p, ws.err_flag);
                else
                    k_os_pass<KT, false><<<tiles, OS_THREADS, smem, st>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, plan.shift[p], plan.bits[p],
                                                                          ws.ghist + (size_t)p * RMAX, ws.status + soff, counters + p, ws.err_flag);
                prof.end();
                soff += (size_t)tiles << plan.bits[p];
                cur ^= 1;
            }
            return cur;
        }
    }
    const RsPlan pl = rs_plan(n, key_bits, ws);
    const Chunking ck = pl.ck;
    const bool big = pl.big;
    int cur = 0, shift = 0;
    for (int p = 0; p < pl.passes; ++p) {
        const int bits = pl.base + (p < pl.rem ? 1 : 0);
        const int R = 1 << bits;
        if (!(p == 0 && ws.first_hist_done)) {
            prof.begin(kid_base);
            k_rs_hist<KT><<<ck.grid, RS_THREADS, R * sizeof(unsigned), st>>>(k[cur], n, ck.chunk, shift, bits, ws.hist);
            prof.end();
        }
        prof.begin(kid_base + 1);
        k_rs_scan<<<(R + 7) / 8, 256, 0, st>>>(ws.hist, ws.prefix, ws.tot, R, ck.grid);
        prof.end();
        const size_t smem = rs_scatter_smem_bytes(bits, sizeof(KT), big ? RS_ITEMS_BIG : RS_ITEMS, big ? RS_BIG_THREADS : RS_THREADS);
        prof.begin(kid_base + 2);
        if (p == 0 && iota_first) {
            if (big) k_rs_scatter<KT, true, RS_ITEMS_BIG, RS_BIG_THREADS><<<ck.grid, RS_BIG_THREADS, smem, st>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, ck.chunk, shift, bits, ws.prefix, ws.tot);
            else k_rs_scatter<KT, true><<<ck.grid, RS_THREADS, smem, st>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, ck.chunk, shift, bits, ws.prefix, ws.tot);
        } else {
            if (big) k_rs_scatter<KT, false, RS_ITEMS_BIG, RS_BIG_THREADS><<<ck.grid, RS_BIG_THREADS, smem, st>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, ck.chunk, shift, bits, ws.prefix, ws.tot);
            else k_rs_scatter<KT, false><<<ck.grid, RS_THREADS, smem, st>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], n, ck.chunk, shift, bits, ws.prefix, ws.tot);
        }
        prof.end();
        cur ^= 1;
        shift += bits;
    }
    return cur;
}

}  // namespace mot
