// remove_static.cuh -- K0: ObstacleTrack::removeStatic (reference MOT.cpp:664-706) as a bit lookup + stable
// compaction, fused with the bounding-box reduction the voxel grid needs.
//
// The reference tests, per point, up to (2t+1)^2 cells of an Eigen::MatrixXd copy of the occupancy grid and
// recomputes atan2/cos/sin of the (frame constant) map yaw.  Here the map is reduced ONCE (mot_set_map) to a
// dilated bitmap  blocked[r][c] = OR over the window of (occ > 50 || occ == -1 || outside the map), so a point
// costs one bit test; cos(-yaw), sin(-yaw) are computed on the host with the same libm float overloads the
// reference calls and enter the kernel as constants.  The index arithmetic keeps the reference's exact fp32
// operation sequence (no FMA) and its truncation toward zero.
#pragma once
#include "common.cuh"

namespace mot {

struct MapParams {
    double origin_x, origin_y;
    float cs, sn;        // cosf(-yaw), sinf(-yaw) from the host's libm
    float resolution;
    int width, height;
    int n_words;         // bitmap words (row-major bit index r*width + c)
};

// blocked bit for every cell (MOT.cpp:681-702 folded over the window; out-of-map counts as unknown).
__global__ void k_build_blocked_bitmap(const int8_t* __restrict__ occ, int W, int H, int t, uint32_t* __restrict__ bits) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= W * H) return;
    const int r = idx / W, c = idx % W;
    bool blocked = false;
    for (int i = -t; i <= t && !blocked; ++i)
        for (int j = -t; j <= t; ++j) {
            const int rr = r + i, cc = c + j;
            if (rr < 0 || cc < 0 || rr >= H || cc >= W) { blocked = true; break; }
            const int v = occ[rr * W + cc];
            if (v > 50 || v == -1) { blocked = true; break; }
        }
    if (blocked) atomicOr(&bits[idx >> 5], 1u << (idx & 31));
}

// keep decision for one point, bit-for-bit the reference's arithmetic (MOT.cpp:674-678).
__device__ __forceinline__ bool rs_keep(const float4& p, const MapParams& mp, const uint32_t* bits) {
    const float x_map = __double2float_rn(__dsub_rn((double)p.x, mp.origin_x));
    const float y_map = __double2float_rn(__dsub_rn((double)p.y, mp.origin_y));
    const float fc = __fdiv_rn(__fsub_rn(__fmul_rn(mp.cs, x_map), __fmul_rn(mp.sn, y_map)), mp.resolution);
    const float fr = __fdiv_rn(__fadd_rn(__fmul_rn(mp.sn, x_map), __fmul_rn(mp.cs, y_map)), mp.resolution);
    if (!(fabsf(fc) < 1073741824.0f) || !(fabsf(fr) < 1073741824.0f)) return false;  // NaN / huge -> drop
    const int col = __float2int_rz(fc), row = __float2int_rz(fr);
    if (row < 0 || col < 0 || row >= mp.height || col >= mp.width) return false;
    const int idx = row * mp.width + col;
    return ((bits[idx >> 5] >> (idx & 31)) & 1u) == 0u;
}

constexpr int RSK_THREADS = 256;
constexpr int RSK_MAX_GRID = 592;
constexpr int RSK_SMEM_BITMAP_MAX = 96 * 1024;  // bitmaps up to 96 KB are staged in shared memory with one TMA bulk copy

struct FrameHeader;  // fwd (api)

// Stages the bitmap into shared memory with a single TMA bulk copy (cp.async.bulk -> UBLKCP) when it fits;
// returns the pointer the lookups should use.
__device__ __forceinline__ const uint32_t* rs_stage_bitmap(const uint32_t* __restrict__ gbits, int n_words, bool use_smem,
                                                           uint32_t* sbits, uint64_t* bar) {
    if (!use_smem) return gbits;
    const uint32_t bytes = (uint32_t)(((n_words * 4) + 15) & ~15);  // buffer is padded to 16 B on the host
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        tma_load_1d(sbits, gbits, bytes, bar);
    }
    const bool ok = mbar_wait_bounded(bar, 0);
    return ok ? sbits : gbits;  // a timed-out copy degrades to global lookups, never to a hang
}

// Pass 1: per-block kept count + bbox of kept points (ordered-int atomics) + non-finite flag.
__global__ void __launch_bounds__(RSK_THREADS) k_rs_count(const float4* __restrict__ pts, int n, int chunk, MapParams mp,
                                                           const uint32_t* __restrict__ gbits, int use_smem,
                                                           int* __restrict__ block_counts, int* __restrict__ bbox /* 6 ordered ints + flag */) {
    extern __shared__ __align__(16) uint32_t rs_sbits[];
    __shared__ uint64_t bar;
    __shared__ int red[33];
    __shared__ int sbox[6];
    const uint32_t* bits = rs_stage_bitmap(gbits, mp.n_words, use_smem != 0, rs_sbits, &bar);
    if (threadIdx.x < 3) sbox[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) sbox[threadIdx.x] = (int)0x80000000;
    __syncthreads();
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    int cnt = 0;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool bad = false;
    for (int i = begin + threadIdx.x; i < end; i += RSK_THREADS) {
        const float4 p = ld_stream(pts + i);
        if (rs_keep(p, mp, bits)) {
            ++cnt;
            if (!(fabsf(p.x) < INFINITY) || !(fabsf(p.y) < INFINITY) || !(fabsf(p.z) < INFINITY)) bad = true;
            mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
            mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(kFull, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(kFull, mx[d], o));
        }
    if (lane_id() == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&sbox[d], float_to_ordered(mn[d]));
            atomicMax(&sbox[3 + d], float_to_ordered(mx[d]));
        }
    }
    if (__any_sync(kFull, bad) && lane_id() == 0) atomicOr(&bbox[6], 1);
    cnt = warp_sum(cnt);
    if (lane_id() == 0) red[warp_id()] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < RSK_THREADS / 32; ++w) s += red[w];
        block_counts[blockIdx.x] = s;
        if (s > 0) {
            for (int d = 0; d < 3; ++d) {
                atomicMin(&bbox[d], sbox[d]);
                atomicMax(&bbox[3 + d], sbox[3 + d]);
            }
        }
    }
}

// Pass 2: stable compaction.  Block base = sum of the preceding blocks' counts; inside a tile the kept points
// are ranked with ballots so the input order is preserved (MOT.cpp:694 appends in input order).
__global__ void __launch_bounds__(RSK_THREADS) k_rs_compact(const float4* __restrict__ pts, int n, int chunk, MapParams mp,
                                                             const uint32_t* __restrict__ gbits, int use_smem,
                                                             const int* __restrict__ block_counts, float4* __restrict__ out,
                                                             int* __restrict__ total_out) {
    extern __shared__ __align__(16) uint32_t rs_sbits[];
    __shared__ uint64_t bar;
    __shared__ int red[36];
    __shared__ int wbase[RSK_THREADS / 32 + 1];
    const uint32_t* bits = rs_stage_bitmap(gbits, mp.n_words, use_smem != 0, rs_sbits, &bar);
    int base = block_prefix_of(block_counts, blockIdx.x, red);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = base + block_counts[blockIdx.x];
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    const int lane = lane_id(), w = warp_id();
    for (int tb = begin; tb < end; tb += RSK_THREADS) {
        const int i = tb + threadIdx.x;
        float4 p = make_float4(0, 0, 0, 0);
        bool keep = false;
        if (i < end) {
            p = ld_stream(pts + i);
            keep = rs_keep(p, mp, bits);
        }
        const unsigned bal = __ballot_sync(kFull, keep);
        __syncthreads();
        if (lane == 0) wbase[w] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int ww = 0; ww < RSK_THREADS / 32; ++ww) {
            const int c = wbase[ww];
            if (ww < w) off += c;
            tot += c;
        }
        if (keep) st_stream(out + base + off + __popc(bal & lanemask_lt()), p);
        base += tot;
    }
}

// ---- single-pass stable compaction (decoupled look-back) ---------------------------------------------------------------------
// k_rs_count + k_rs_compact read the cloud twice (32 N + 16 M bytes).  k_compact_onepass reads it once (16 N + 16 M): a CTA takes
// the next tile (ticket: every predecessor tile is running or done), decides its points, publishes the tile's kept count in
// one 32-bit status word (2 flag bits + 30 value bits, so no fence is needed), obtains its exclusive prefix by walking back
// over the predecessors' words 32 at a time, and writes its kept points in input order (MOT.cpp:694 appends in input order).
// The same pass accumulates the bounding box / non-finite flag of the kept points (the voxel grid needs them) and, for a
// batch of frames, the frame boundaries of the compacted cloud.
// MODE 0: keep = removeStatic's map lookup (MOT.cpp:664-706); MODE 1: keep = all three coordinates finite
// (pcl::removeNaNFromPointCloud, used behind fromROSMsg).
constexpr int C1P_THREADS = 256;
constexpr int C1P_ITEMS = 8;
constexpr int C1P_TILE = C1P_THREADS * C1P_ITEMS;
constexpr unsigned C1P_AGG = 1u << 30, C1P_INCL = 2u << 30, C1P_VALUE = (1u << 30) - 1u;
constexpr int C1P_SPIN_LIMIT = 1 << 24;  // a stuck look-back raises flag 2 instead of hanging the GPU

// warp 0 of a tile: exclusive prefix of the tile's count over all preceding tiles (the tile's own aggregate is already
// published); afterwards the tile's word holds its inclusive prefix.  Returns the prefix in every lane.
__device__ __forceinline__ int c1p_lookback(unsigned* status, int tile, int tile_total, int* err_flag) {
    int prefix = 0;
    const int lane = lane_id();
    int t = tile - 1, spins = 0;
    while (t >= 0) {
        const int mine = t - lane;
        const unsigned sv = mine >= 0 ? __ldcg(status + mine) : C1P_INCL;  // before tile 0: an inclusive prefix of 0
        const unsigned f = sv >> 30;
        const unsigned incl_m = __ballot_sync(kFull, f == 2u), empty_m = __ballot_sync(kFull, f == 0u);
        const int first_incl = incl_m ? __ffs(incl_m) - 1 : 32;                              // nearest predecessor holding an inclusive prefix
        const unsigned need = first_incl >= 31 ? 0xffffffffu : ((2u << first_incl) - 1u);     // lanes 0 .. first_incl
        if (empty_m & need) {
            if (++spins > C1P_SPIN_LIMIT) { if (lane == 0) atomicOr(err_flag, 2); break; }
            continue;
        }
        int v = ((need >> lane) & 1u) ? (int)(sv & C1P_VALUE) : 0;
        v = warp_sum(v);
        prefix += v;
        if (incl_m) break;
        t -= 32;
    }
    if (lane == 0) __stcg(status + tile, C1P_INCL | (unsigned)(prefix + tile_total));
    return prefix;
}

// The same pass over 32-bit keys in index order: keeps (key, index) of every key != drop.  Used to shrink the CSR sort to the
// points of the clusters that survive the size filter (cluster_table.cuh).
template <int ITEMS>
__global__ void __launch_bounds__(C1P_THREADS) k_compact_keys_onepass(const uint32_t* __restrict__ keys, int n, uint32_t drop, uint32_t* __restrict__ kout,
                                                                       uint32_t* __restrict__ vout, unsigned* status, int n_tiles,
                                                                       int* __restrict__ err_flag) {
    static_assert(ITEMS % 4 == 0, "16-byte loads");
    constexpr int TILE = C1P_THREADS * ITEMS;
    __shared__ int scratch[36];
    __shared__ int s_tile, s_base;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(status + n_tiles, 1u);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) return;
    const int i0 = tile * TILE + threadIdx.x * ITEMS;
    uint32_t k[ITEMS];
    unsigned keepm = 0;
    if (i0 + ITEMS <= n) {  // 16-byte loads (i0 is a multiple of ITEMS)
#pragma unroll
        for (int v = 0; v < ITEMS / 4; ++v) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(keys + i0) + v);
            k[4 * v] = a.x; k[4 * v + 1] = a.y; k[4 * v + 2] = a.z; k[4 * v + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) k[j] = i0 + j < n ? keys[i0 + j] : drop;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) keepm |= (k[j] != drop ? 1u : 0u) << j;
    int tile_total;
    const int excl = block_exclusive_scan(__popc(keepm), scratch, &tile_total);
    if (threadIdx.x == 0) __stcg(status + tile, (tile == 0 ? C1P_INCL : C1P_AGG) | (unsigned)tile_total);
    if (warp_id() == 0) {
        const int prefix = c1p_lookback(status, tile, tile_total, err_flag);
        if (lane_id() == 0) s_base = prefix;
    }
    // kept pairs go through shared memory so that the tile leaves as one coalesced run (written straight from the
    // registers every lane of a store hits its own sector: 106 us for 16.8 M keys; staged: see profiles/r02_csr_compact.txt)
    __shared__ uint32_t s_k[TILE], s_v[TILE];
    int w = excl;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
        if ((keepm >> j) & 1u) {
            s_k[w] = k[j];
            s_v[w] = (uint32_t)(i0 + j);
            ++w;
        }
    __syncthreads();
    const int base = s_base;
    for (int t = threadIdx.x; t < tile_total; t += C1P_THREADS) {
        kout[base + t] = s_k[t];
        vout[base + t] = s_v[t];
    }
}

template <int MODE>
__global__ void __launch_bounds__(C1P_THREADS) k_compact_onepass(const float4* __restrict__ pts, int n, MapParams mp, const uint32_t* __restrict__ gbits,
                                                                  int use_smem, float4* __restrict__ out, unsigned* status /* [tiles] + ticket */,
                                                                  int n_tiles, int* __restrict__ total_out, int* __restrict__ bbox,
                                                                  const int* __restrict__ frame_off_in, int n_frames, int* __restrict__ frame_off_out,
                                                                  int* __restrict__ err_flag) {
    extern __shared__ __align__(16) uint32_t rs_sbits[];
    __shared__ uint64_t bar;
    __shared__ int scratch[36];
    __shared__ int s_tile, s_base;
    __shared__ int sbox[6];
    const uint32_t* bits = MODE == 0 ? rs_stage_bitmap(gbits, mp.n_words, use_smem != 0, rs_sbits, &bar) : nullptr;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(status + n_tiles, 1u);
    if (threadIdx.x < 3) sbox[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) sbox[threadIdx.x] = (int)0x80000000;
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) return;
    const int tile_begin = tile * C1P_TILE;
    // blocked arrangement: a thread owns C1P_ITEMS consecutive points (one 128-byte line), so ranks are input order
    const int i0 = tile_begin + threadIdx.x * C1P_ITEMS;
    float4 p[C1P_ITEMS];
    unsigned keepm = 0;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool bad = false;
#pragma unroll
    for (int k = 0; k < C1P_ITEMS; ++k)
        if (i0 + k < n) p[k] = ld_stream(pts + i0 + k);
#pragma unroll
    for (int k = 0; k < C1P_ITEMS; ++k) {
        if (i0 + k >= n) break;
        const bool fin = (fabsf(p[k].x) < INFINITY) && (fabsf(p[k].y) < INFINITY) && (fabsf(p[k].z) < INFINITY);
        const bool keep = MODE == 0 ? rs_keep(p[k], mp, bits) : fin;
        if (keep) {
            keepm |= 1u << k;
            bad |= !fin;
            mn[0] = fminf(mn[0], p[k].x); mn[1] = fminf(mn[1], p[k].y); mn[2] = fminf(mn[2], p[k].z);
            mx[0] = fmaxf(mx[0], p[k].x); mx[1] = fmaxf(mx[1], p[k].y); mx[2] = fmaxf(mx[2], p[k].z);
        }
    }
    int tile_total;
    const int excl = block_exclusive_scan(__popc(keepm), scratch, &tile_total);
    // publish, then look back (warp 0)
    if (threadIdx.x == 0) __stcg(status + tile, (tile == 0 ? C1P_INCL : C1P_AGG) | (unsigned)tile_total);
    if (warp_id() == 0) {
        const int prefix = c1p_lookback(status, tile, tile_total, err_flag);
        if (lane_id() == 0) {
            s_base = prefix;
            if (tile == n_tiles - 1) *total_out = prefix + tile_total;
        }
    }
    // bbox of the kept points while the look-back runs
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(kFull, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(kFull, mx[d], o));
        }
    if (lane_id() == 0 && tile_total > 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&sbox[d], float_to_ordered(mn[d]));
            atomicMax(&sbox[3 + d], float_to_ordered(mx[d]));
        }
    }
    if (__any_sync(kFull, bad) && lane_id() == 0) atomicOr(&bbox[6], 1);
    __syncthreads();
    const int base = s_base;
    if (threadIdx.x < 6 && tile_total > 0) {
        if (threadIdx.x < 3) atomicMin(&bbox[threadIdx.x], sbox[threadIdx.x]);
        else atomicMax(&bbox[threadIdx.x], sbox[threadIdx.x]);
    }
    int w = base + excl;
#pragma unroll
    for (int k = 0; k < C1P_ITEMS; ++k)
        if ((keepm >> k) & 1u) st_stream(out + w++, p[k]);
    // frame boundaries of the compacted cloud: boundary b (an input index) moves to the output position point b gets / would get
    if (frame_off_out) {
        const int tile_end = min(n, tile_begin + C1P_TILE);
        int lo = 0, hi = n_frames + 1;  // first f with frame_off_in[f] >= tile_begin
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (frame_off_in[mid] < tile_begin) lo = mid + 1; else hi = mid;
        }
        for (int f = lo; f <= n_frames; ++f) {
            const int b = frame_off_in[f];
            if (b >= tile_end && !(b == n && tile == n_tiles - 1)) break;
            if (b == n) { if (threadIdx.x == 0) frame_off_out[f] = base + tile_total; continue; }
            if (b >= i0 && b < i0 + C1P_ITEMS) frame_off_out[f] = base + excl + __popc(keepm & ((1u << (b - i0)) - 1u));
        }
    }
}

// Bounding box + finiteness of an already compacted cloud (mot_cluster without removeStatic).  Two points per
// thread and iteration in flight, block-level reduction in shared memory, six global atomics per block.
__global__ void __launch_bounds__(256) k_bbox(const float4* __restrict__ pts, int n, int* __restrict__ bbox) {
    __shared__ float sred[6][8];
    __shared__ int sbad;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool bad = false;
    if (threadIdx.x == 0) sbad = 0;
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {
        const float4 p = ld_stream(pts + i), q = ld_stream(pts + i + stride);
        bad |= !(fabsf(p.x) < INFINITY) || !(fabsf(p.y) < INFINITY) || !(fabsf(p.z) < INFINITY);
        bad |= !(fabsf(q.x) < INFINITY) || !(fabsf(q.y) < INFINITY) || !(fabsf(q.z) < INFINITY);
        mn[0] = fminf(mn[0], fminf(p.x, q.x)); mn[1] = fminf(mn[1], fminf(p.y, q.y)); mn[2] = fminf(mn[2], fminf(p.z, q.z));
        mx[0] = fmaxf(mx[0], fmaxf(p.x, q.x)); mx[1] = fmaxf(mx[1], fmaxf(p.y, q.y)); mx[2] = fmaxf(mx[2], fmaxf(p.z, q.z));
    }
    if (i < n) {
        const float4 p = ld_stream(pts + i);
        bad |= !(fabsf(p.x) < INFINITY) || !(fabsf(p.y) < INFINITY) || !(fabsf(p.z) < INFINITY);
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(kFull, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(kFull, mx[d], o));
        }
    __syncthreads();
    if (lane_id() == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { sred[d][warp_id()] = mn[d]; sred[3 + d][warp_id()] = mx[d]; }
    }
    if (__any_sync(kFull, bad) && lane_id() == 0) atomicOr(&sbad, 1);
    __syncthreads();
    if (threadIdx.x < 6) {
        const int d = threadIdx.x;
        float v = sred[d][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = d < 3 ? fminf(v, sred[d][w]) : fmaxf(v, sred[d][w]);
        if (d < 3) atomicMin(&bbox[d], float_to_ordered(v));
        else atomicMax(&bbox[d], float_to_ordered(v));
    }
    if (threadIdx.x == 0 && sbad) atomicOr(&bbox[6], 1);
}

// ---- sensor_msgs/PointCloud2 wire format -> pcl::PointXYZ (what pcl::fromROSMsg does at MOT.cpp:448-449) ----
// One thread per point: x, y, z FLOAT32 fields at byte offsets off[0..2] inside a point_step-byte record (any alignment,
// either endianness).  Pass 1 unpacks to float4 (w = 1) and counts the finite points per block; pass 2 is the same stable
// compaction as removeStatic when non-finite points are to be dropped.
__device__ __forceinline__ float pc2_load_f32(const uint8_t* p, bool aligned, bool bigendian) {
    uint32_t v;
    if (aligned) v = *reinterpret_cast<const uint32_t*>(p);
    else v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    if (bigendian) v = __byte_perm(v, 0, 0x0123);
    return __uint_as_float(v);
}
__global__ void __launch_bounds__(256) k_pc2_unpack(const uint8_t* __restrict__ data, int n, int chunk, uint32_t point_step, uint32_t off_x,
                                                     uint32_t off_y, uint32_t off_z, int bigendian, float4* __restrict__ out,
                                                     int* __restrict__ block_counts) {
    __shared__ int red[8];
    const bool aligned = ((point_step | off_x | off_y | off_z) & 3u) == 0u;
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    int cnt = 0;
    for (int i = begin + threadIdx.x; i < end; i += 256) {
        const uint8_t* rec = data + (size_t)i * point_step;
        float4 p;
        p.x = pc2_load_f32(rec + off_x, aligned, bigendian != 0);
        p.y = pc2_load_f32(rec + off_y, aligned, bigendian != 0);
        p.z = pc2_load_f32(rec + off_z, aligned, bigendian != 0);
        p.w = 1.0f;
        out[i] = p;
        cnt += (fabsf(p.x) < INFINITY) && (fabsf(p.y) < INFINITY) && (fabsf(p.z) < INFINITY);
    }
    cnt = warp_sum(cnt);
    if (lane_id() == 0) red[warp_id()] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        block_counts[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) k_pc2_compact(const float4* __restrict__ in, int n, int chunk, const int* __restrict__ block_counts,
                                                      float4* __restrict__ out, int* __restrict__ total_out) {
    __shared__ int red[36];
    __shared__ int wbase[9];
    int base = block_prefix_of(block_counts, blockIdx.x, red);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = base + block_counts[blockIdx.x];
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    const int lane = lane_id(), w = warp_id();
    for (int tb = begin; tb < end; tb += 256) {
        const int i = tb + threadIdx.x;
        float4 p = make_float4(0, 0, 0, 0);
        bool keep = false;
        if (i < end) {
            p = in[i];
            keep = (fabsf(p.x) < INFINITY) && (fabsf(p.y) < INFINITY) && (fabsf(p.z) < INFINITY);
        }
        const unsigned bal = __ballot_sync(kFull, keep);
        __syncthreads();
        if (lane == 0) wbase[w] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) {
            const int c = wbase[ww];
            if (ww < w) off += c;
            tot += c;
        }
        if (keep) out[base + off + __popc(bal & lanemask_lt())] = p;
        base += tot;
    }
}

}  // namespace mot
