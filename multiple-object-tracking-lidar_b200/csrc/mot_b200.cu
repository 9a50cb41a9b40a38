// mot_b200.cu -- host side of libmot_b200.so: the C ABI declared in include/mot_b200.h and the per-frame
// launch sequence.  C++ host code calling hand-written sm_100a kernels; no PyTorch, no CPU fallback.
//
// Frame pipeline (one CUDA stream per handle -- plus a low-priority side stream for the K4 kernels -- and two host round
// trips of 32 bytes each):
//   K0  k_rs_count / k_rs_compact     removeStatic: bit lookup + stable compaction (+ bbox of kept points)
//   --  S1: read M and the bounding box (sizes the grid: key width, radix passes, hash table)
//   K1  k_cell_keys                   fp64 cell coordinates -> voxel key
//   K2  k_rs_hist/scan/scatter        LSD radix sort of (key, index), ceil(bits/10) passes
//   K3  k_cells_count / k_hash_clear / k_cells_write   reorder to sorted SoA, fine/coarse cell tables, coarse-cell hash
//   K4  k_coarse_records              per coarse cell: record + resolved forward half stencil (13 neighbours)
//       k_uf_sparse                   warp per coarse cell: TMA-staged neighbourhood, exact brute-force sweep, atomicMin hooking
//       k_uf_dense<1>, <2>            neighbourhoods too large for the tile: witness search per fine-cell pair (rings 1, 2)
//       (MOT_UF_MODE=0: k_uf_pairs<1>, <2>, the first-generation warp-per-fine-cell kernels)
//   K5  k_uf_flatten                  pointer jumping (in place between the dense rings, then into root[])
//   K6  k_comp_accumulate/k_kept_list component sizes, [min,max] filter
//   --  S2: read K (sizes the cluster sort and the CSR partition)
//       k_clusters_small | radix64    order clusters (size desc, min index asc), CSR offsets
//       k_point_rank + radix32        stable partition of point indices by cluster rank = CSR indices
//   K7  k_stats_init/accumulate/finalize  point-parallel segmented reduction: count / mean / bbox
//   K8  k_farthest_pair/k_circumcentre  the reference's getCentroid
//   K9  k_ihgp_step                   batched track filter (separate entry point)
#include <sched.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mot_b200.h"
#include "cell_uf.cuh"
#include "cluster_table.cuh"
#include "common.cuh"
#include "frame_small.cuh"
#include "grid_uf.cuh"
#include "ihgp.cuh"
#include "radix_sort.cuh"
#include "remove_static.cuh"
#include "tracks.cuh"

using namespace mot;

static_assert(sizeof(mot_cluster_stat) == sizeof(ClusterStat), "stat layout");
static_assert(sizeof(mot_cluster_stat) == 40, "stat layout");
static_assert(sizeof(mot_obstacle) == sizeof(ObstacleRow) && sizeof(mot_obstacle) == 48, "obstacle layout");

// kernel ids for the launch counter / per-kernel profile (mot_profile_read)
enum KernelId {
    KID_RS_COUNT, KID_RS_COMPACT, KID_BBOX, KID_KEYS, KID_SORT_HIST, KID_SORT_SCAN, KID_SORT_SCATTER, KID_CELLS_COUNT, KID_HASH_CLEAR, KID_CELLS_WRITE,
    KID_UF1, KID_FLATTEN1, KID_UF2, KID_FLATTEN2, KID_COARSE_REC, KID_UF_COARSE, KID_UF_DENSE1, KID_UF_DENSE, KID_COMP_ACC, KID_KEPT_LIST, KID_CL_SMALL, KID_CLSORT_HIST, KID_CLSORT_SCAN,
    KID_CLSORT_SCATTER, KID_CL_COUNT, KID_CL_FINALIZE, KID_FRAME_CL_OFF, KID_POINT_RANK, KID_PART_HIST, KID_PART_SCAN, KID_PART_SCATTER,
    KID_LOCALIZE, KID_STATS_INIT, KID_STATS, KID_STATS_FIN, KID_FARTHEST_PAIR, KID_CIRCUMCENTRE, KID_IHGP, KID_VOX_KEYS, KID_VOX_HIST, KID_VOX_SCAN, KID_VOX_SCATTER, KID_SEG_COUNT, KID_SEG_WRITE, KID_VOX_FIN, KID_PC2_UNPACK, KID_PC2_COMPACT, KID_ASSOCIATE, KID_TRACKS_PURGE, KID_UF_PAIR, KID_CELL_LOCAL, KID_UF_CROSS, KID_UF_HEAVY1, KID_FLATTEN_IF, KID_UF_HEAVY2, KID_UF_SURV, KID_UF_WALK, KID_UF_FUSED, KID_CSR_COMPACT, KID_CELL_LOCAL_DENSE, KID_HASH_BUILD, KID_FS_FRONT, KID_FS_EDGES, KID_FS_TABLES, KID_FS_FARTHEST, KID_FS_FINISH, KID_TRACKS_APPLY, KID_N
};
static const char* const kKernelNames[KID_N] = {
    "k_rs_count", "k_compact_onepass<map>", "k_bbox", "k_cell_keys", "k_rs_hist[cells]", "k_rs_scan[cells]", "k_rs_scatter[cells]", "k_cells_count",
    "k_hash_clear", "k_cells_write", "k_uf_pairs<1>", "k_uf_flatten<in-place>", "k_uf_pairs<2>", "k_uf_flatten<root>", "k_coarse_records", "k_uf_sparse", "k_uf_dense<1>", "k_uf_dense<2>", "k_comp_accumulate", "k_kept_list",
    "k_clusters_small", "k_rs_hist[clusters]", "k_rs_scan[clusters]", "k_rs_scatter[clusters]", "k_clusters_count", "k_clusters_finalize",
    "k_frame_cluster_offsets", "k_point_rank", "k_rs_hist[csr]", "k_rs_scan[csr]", "k_rs_scatter[csr]", "k_localize_indices",
    "k_stats_init", "k_stats_accumulate", "k_stats_finalize", "k_farthest_pair", "k_circumcentre", "k_ihgp_step", "k_voxel_keys", "k_rs_hist[voxel]", "k_rs_scan[voxel]", "k_rs_scatter[voxel]",
    "k_seg_count", "k_seg_write", "k_voxel_finalize", "k_pc2_unpack", "k_compact_onepass<finite>", "k_associate", "k_tracks_purge", "k_uf_sparse2", "k_cell_local", "k_uf_cross", "k_uf_heavy<1>", "k_uf_flatten_if", "k_uf_heavy<2>", "k_uf_survivors", "k_uf_walk", "k_uf_fused", "k_compact_keys_onepass", "k_cell_local_dense", "k_hash_build", "k_fs_front", "k_fs_edges", "k_fs_tables", "k_fs_farthest", "k_fs_finish", "k_tracks_apply"};

struct mot_handle {
    int device = 0;
    size_t max_points = 0, max_tracks = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int num_sms = 148;

    // launch parameters (reference defaults, MOT.cpp:90-92)
    float tol = 0.15f;
    int min_size = 5, max_size = 200;

    // map
    bool have_map = false;
    MapParams mp{};
    uint32_t* d_bits = nullptr;
    size_t bits_capacity_words = 0;

    // frame workspace
    float4 *d_in = nullptr, *d_pts = nullptr, *d_spts = nullptr;
    void* d_keys[2] = {nullptr, nullptr};  // 8 bytes per point each (u32 or u64 keys)
    uint32_t* d_vals[2] = {nullptr, nullptr};
    int *d_fc_start = nullptr, *d_cc_first = nullptr, *d_parent = nullptr, *d_root = nullptr;
    int *d_csize = nullptr, *d_cmin = nullptr, *d_crank = nullptr, *d_labels = nullptr;
    void* d_hkeys = nullptr;
    int* d_hvals = nullptr;
    size_t hash_capacity = 0;
    RadixWorkspace rws;
    int* d_blk = nullptr;     // 32 * 1024 per-block counters (k_cells_*: 4 rows of up to CELLW_MAX_GRID)
    int* d_counts = nullptr;  // CNT_N ints
    cudaEvent_t sync_ev = nullptr;  // blocking-sync event (MOT_SYNC=block): host waits sleep instead of spinning
    int sync_mode = 0;  // 0 spin, 1 yield, 2 block (MOT_SYNC)
    // The union-find kernels (latency bound, they hold every SM for ~2 ms per batch) run on a lowest-priority side stream
    // and everything else on a highest-priority one: when several handles share the GPU, the bandwidth-bound kernels of
    // the other handles are placed on the SMs first instead of queueing behind a union-find launch (MOT_UF_PRIO=0: off).
    cudaStream_t uf_stream = nullptr;
    cudaEvent_t ev_uf[2] = {nullptr, nullptr};
    int uf_pair = 1;           // k_uf_sparse2 (half a warp per coarse cell); MOT_UF_PAIR=0: k_uf_sparse (a warp per cell)
    int uf_blocks_per_sm = 5;  // resident CTAs of k_uf_sparse per SM (MOT_UF_BLOCKS; fewer leaves room for other streams' kernels)
    int* d_bbox = nullptr;    // 8 ints
    uint64_t* d_ckeys[2] = {nullptr, nullptr};
    uint32_t* d_croots[2] = {nullptr, nullptr};
    int* d_cl_offsets = nullptr;
    int* d_frame_offsets = nullptr;   // frame boundaries of the cloud being clustered (batch mode)
    int* d_frame_offsets_in = nullptr;  // ... of the input cloud, before removeStatic
    unsigned* d_c1p_status = nullptr;  // one-pass compaction: tile status words + ticket
    size_t c1p_tiles = 0;
    // speculative grid plan: the grid of the previous call on this handle, padded to the power-of-two range of its key bits
    bool plan_valid = false;
    GridCodec plan_g{};
    int plan_bits = 0;
    float plan_tol = 0.0f;
    int plan_spec = 1;        // MOT_PLAN_SPEC=0: always take the bounding box first (one more kernel and one more host round trip per call)
    int plan_hits = 0, plan_misses = 0;
    // single-launch path for small frames (frame_small.cuh)
    int fs_cluster = 0;            // CTAs of the cluster (16 or 8; 0 = not available)
    int fs_max_points = 131072;    // MOT_SMALL_POINTS: frames of at most this many input points try it (0 = never)
    int fs_cell_cap = 64;          // MOT_SMALL_CELLCAP: a fine cell with more points hands the frame to the general path
    int* d_fs_state = nullptr;
    FsArgs* d_fs_args = nullptr;   // the kernels read their arguments from device memory, so that one instantiated graph serves every call
    FsArgs* h_fs_args = nullptr;   // pinned staging of the arguments (source of the graph's copy node)
    cudaGraphExec_t fs_graph = nullptr;
    int fs_use_graph = 1;          // MOT_SMALL_GRAPH=0: five plain launches instead of one graph launch
    int fs_skip = 0;               // calls left that go straight to the general path after a fallback
    int fs_hits = 0, fs_misses = 0;
    bool tim_stages = true;        // false: the last call recorded only its first and last event (small-frame path)
    unsigned long long* d_fs_phase_ns = nullptr;
    int keys_hist_fused = 0;  // MOT_KEYS_HIST=1: k_cell_keys_hist instead of k_cell_keys + the first k_rs_hist (measured neutral: 107 + 2x35 vs 81 + 3x32 us)
    int csr_compact = 1;      // drop the points of filtered-out components before the CSR sort when they are >= 40 % (c2: 77 %, 367 -> 259 us); MOT_CSR_COMPACT=0: always sort all
    int rs_mode = 1;  // 1: single-pass compaction (decoupled look-back); 0: count + compact (MOT_RS_MODE)
    float* d_frame_stamps = nullptr;  // per-frame centroid intensity (batch mode)
    bool have_frame_stamps = false;
    int* d_frame_cl_offsets = nullptr;
    size_t frame_capacity = 0;
    ClusterStat* d_stats = nullptr;
    StatAcc* d_statacc = nullptr;
    int4* d_crec = nullptr;
    int* d_dense_list = nullptr;
    int* d_nbr = nullptr;
    int dense_cap = 0;
    int uf_mode = 3;   // MOT_UF_MODE 3 (default): 2 for calls with at least uf_auto_points points, else 1 -- the cell path is throughput
                       // oriented (latency-bound kernels that need ~10^5 coarse cells in flight), the sweep has the shorter critical path;
                       // 2: cell boxes + local components + fused lookup / root skip / walk (cell_uf.cuh);
                       // 1: coarse-cell warps + TMA staging + brute-force sweep; 0: v1 two-phase fine-cell warps
    int uf_auto_points = 3 << 19;  // MOT_UF_AUTO: crossover of the two paths (1.5 M points)
    int uf_mode_now = 2;           // what the current call uses
    void* d_ckey = nullptr;      // sorted coarse keys (u32 or u64)
    unsigned char* d_fcode = nullptr;  // child code (0..7) of every fine cell
    float4 *d_cbox = nullptr, *d_fbox = nullptr;  // AABB of every coarse / fine cell (2 x float4 each)
    int2 *d_heavy1 = nullptr, *d_heavy2 = nullptr;
    int heavy_cap = 0;
    int l2_persist_mb = 0;  // MOT_L2_PERSIST: MB of L2 set aside (persisting access window) for parent[] / the coarse records during the union-find kernels
    int l2_persist_what = 0; // MOT_L2_WHAT: 0 = parent[], 1 = coarse records
    int uf_tile_batches = 0;  // MOT_UF_TILE: k_uf_fused walks the cells in tiles of this many 32-cell batches (0 = one tile)
    int uf_row_inner = 0;   // MOT_UF_ROWINNER: k_uf_fused item order (see the kernel)
    int cld_threads = 256;  // CTA width of k_cell_local_dense (MOT_CLD_THREADS: 128 / 256 / 512)
    int cell_dense = 1;     // batches of > 2048 points go to k_cell_local_dense (MOT_CELL_DENSE=0: one warp walks any batch)
    int uf_light = 256;     // fine-cell pairs with at most this many point pairs are searched by one thread (MOT_UF_LIGHT)
    int uf_cross_blocks = 4;  // resident CTAs of k_uf_cross per SM (MOT_UF_XBLOCKS)
    int uf_split = 0;        // MOT_UF_SPLIT=1: face rows and the remaining rows as two launches
    int uf_xmode = 2;        // MOT_UF_XMODE 2: k_uf_fused (default); 1: k_uf_survivors + k_uf_walk per phase; 0: k_uf_cross
    unsigned uf_phases[3] = {0x01u, 0x06u, 0x18u};  // neighbour rows per phase (MOT_UF_PHASES, e.g. "1,6,24")
    int uf_n_phases = 3;
    int2* d_tasks = nullptr;  // surviving cell pairs of one phase (A | code << 27, B)
    int task_cap = 0;
    int uf_walk_blocks = 8;   // resident CTAs of k_uf_walk per SM (MOT_UF_WBLOCKS)
    int uf_fused_blocks = UFF_MIN_BLOCKS;  // resident CTAs of k_uf_fused per SM (MOT_UF_FBLOCKS)
    int uf_tma = 1;
    float4* d_centroids = nullptr;
    PairCand* d_cands = nullptr;
    size_t table_capacity = 0, cand_capacity = 0;
    int* h_pinned = nullptr;  // 32 ints: counts + bbox readback
    int* h_pinned_fo = nullptr;  // pinned staging of a batch's frame offsets (+ stamps)
    // small transfers between PAGEABLE caller buffers and the device go through pinned staging: a pageable cudaMemcpyAsync is a
    // synchronous driver call (~20-60 us each; the one-frame host entry makes six of them), pinned copies are queued in ~2 us
    unsigned char* h_stage = nullptr;  // [STAGE_IN + STAGE_OUT]
    int stage_on = 1;                  // MOT_HOST_STAGE=0: every copy goes straight to / from the caller's pointer

    // IHGP
    bool ihgp_ready = false;
    double ihgp_consts[2][16];
    IhgpAxis ihgp_axis[2];
    double ihgp_dt = 0.1;
    float ihgp_tau = 0.01f;
    int ihgp_L = 10;
    float4* d_rings = nullptr;
    double* d_mstate = nullptr;
    float4* d_posvel = nullptr;
    int* d_track_ids = nullptr;
    ObstacleRow* d_obstacles = nullptr;
    uint8_t* d_raw = nullptr;
    // on-device track table: ping-pong halves for the purge
    int* d_trk_ids[2] = {nullptr, nullptr};
    float4* d_trk_rings[2] = {nullptr, nullptr};
    double* d_trk_m[2] = {nullptr, nullptr};
    int *d_trk_meta = nullptr, *d_trk_seen = nullptr, *d_ent_ids = nullptr, *d_ent_slot = nullptr, *d_ent_occ = nullptr, *d_ent_next = nullptr;
    float4* d_ent_prev = nullptr;  // last observation an entry replaced (k_associate_fast -> k_tracks_apply)
    int assoc_fast = 1;            // MOT_ASSOC_FAST=0: the one-kernel association (k_associate) for every table size
    float4* d_centroids_in = nullptr;
    int trk_cur = 0, trk_L = 0, trk_spin = 0, trk_n = 0;
    bool trk_first = true;
    size_t raw_capacity = 0;
    size_t ring_capacity = 0;

    // last result
    bool have_result = false;
    const float4* res_cloud = nullptr;  // the clustered cloud on the device (d_pts or the caller's pointer)
    int res_M = 0, res_K = 0, res_total = 0, res_idx_buf = 0, res_frames = 1;
    int res_fine = 0, res_coarse = 0, res_key_bits = 0;
    int res_counters[CNT_N] = {};
    int sorted_buf = 0;
    bool res_centroids = false;
    Prof prof;
    float prof_ms[KID_N] = {};
    int prof_n[KID_N] = {};
    cudaEvent_t ev[6] = {};
    cudaEvent_t timer_ev[2] = {};
    mot_timings tim{};
};

// every kernel launch goes through LAUNCH: counts it and, when profiling is on, brackets it with an event pair
// Host wait for the handle's stream.  Default: cudaStreamSynchronize (spins, lowest latency).  MOT_SYNC=yield polls an
// event and gives the core away between polls -- for hosts that run more handles than cores (several ranks x several
// streams per GPU on one node); MOT_SYNC=block sleeps on a cudaEventBlockingSync event (slowest to wake up).
static inline cudaError_t mot_sync(mot_handle* h) {
    if (h->sync_mode == 0) return cudaStreamSynchronize(h->stream);
    cudaError_t e = cudaEventRecord(h->sync_ev, h->stream);
    if (e != cudaSuccess) return e;
    if (h->sync_mode == 2) return cudaEventSynchronize(h->sync_ev);
    while ((e = cudaEventQuery(h->sync_ev)) == cudaErrorNotReady) sched_yield();
    return e;
}

#define LAUNCH(kid, ...)        \
    do {                        \
        h->prof.begin(kid);     \
        __VA_ARGS__;            \
        h->prof.end();          \
    } while (0)

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return MOT_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int fail(mot_handle* h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}

int ceil_log2(long long v) {  // bits needed to represent values 0 .. v-1
    int b = 0;
    while ((1ll << b) < v) ++b;
    return b;
}

template <typename T>
cudaError_t dalloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T) + 256); }

int ensure_tables(mot_handle* h, size_t K, size_t cands) {
    if (K > h->table_capacity) {
        size_t cap = K + K / 2 + 1024;
        if (h->d_stats) cudaFree(h->d_stats);
        if (h->d_centroids) cudaFree(h->d_centroids);
        if (h->d_statacc) cudaFree(h->d_statacc);
        h->d_stats = nullptr; h->d_centroids = nullptr; h->d_statacc = nullptr;
        CK(dalloc(&h->d_stats, cap));
        CK(dalloc(&h->d_statacc, cap));
        CK(dalloc(&h->d_centroids, cap));
        h->table_capacity = cap;
    }
    if (cands > h->cand_capacity) {
        size_t cap = cands + cands / 2 + 1024;
        if (h->d_cands) cudaFree(h->d_cands);
        h->d_cands = nullptr;
        CK(dalloc(&h->d_cands, cap));
        h->cand_capacity = cap;
    }
    return MOT_OK;
}

// ------------------------------------------------------------------------------------------------------------
// removeStatic stage: d_src[n] -> d_pts[M], M and bbox left in device memory (read at S1 by cluster_core)
// ------------------------------------------------------------------------------------------------------------
// d_src[n] -> d_dst[M] (stable); M is left in d_counts[CNT_M], the bbox of the kept points in d_bbox.  In batch mode
// (n_frames > 1) the frame boundaries d_frame_offsets_in are mapped to the compacted cloud (d_frame_offsets).
int enqueue_remove_static(mot_handle* h, const float4* d_src, int n, float4* d_dst, int n_frames = 1) {
    const size_t bitmap_bytes = (size_t)((h->mp.n_words * 4 + 15) & ~15);
    const int use_smem = bitmap_bytes <= (size_t)RSK_SMEM_BITMAP_MAX ? 1 : 0;
    const size_t smem = use_smem ? bitmap_bytes : 0;
    if (h->rs_mode == 1) {
        const int tiles = (n + C1P_TILE - 1) / C1P_TILE;
        CK(cudaMemsetAsync(h->d_c1p_status, 0, ((size_t)tiles + 1) * sizeof(unsigned), h->stream));
        LAUNCH(KID_RS_COMPACT, k_compact_onepass<0><<<tiles, C1P_THREADS, smem, h->stream>>>(d_src, n, h->mp, h->d_bits, use_smem, d_dst, h->d_c1p_status, tiles,
                                                                                          h->d_counts + CNT_M, h->d_bbox,
                                                                                          n_frames > 1 ? h->d_frame_offsets_in : nullptr, n_frames,
                                                                                          n_frames > 1 ? h->d_frame_offsets : nullptr,
                                                                                          h->d_counts + CNT_FLAGS));
    } else {
        if (n_frames > 1) return fail(h, MOT_ERR_STATE, "MOT_RS_MODE=0 does not support removeStatic on frame batches");
        const Chunking ck = make_chunking(n, RSK_THREADS, RSK_MAX_GRID);
        LAUNCH(KID_RS_COUNT, k_rs_count<<<ck.grid, RSK_THREADS, smem, h->stream>>>(d_src, n, ck.chunk, h->mp, h->d_bits, use_smem, h->d_blk, h->d_bbox));
        LAUNCH(KID_RS_COMPACT, k_rs_compact<<<ck.grid, RSK_THREADS, smem, h->stream>>>(d_src, n, ck.chunk, h->mp, h->d_bits, use_smem, h->d_blk,
                                                                                      d_dst, h->d_counts + CNT_M));
    }
    CK(cudaGetLastError());
    return MOT_OK;
}

int ensure_raw(mot_handle* h, size_t bytes) {
    if (bytes > h->raw_capacity) {
        if (h->d_raw) cudaFree(h->d_raw);
        h->d_raw = nullptr;
        h->raw_capacity = 0;
        CK(cudaMalloc(reinterpret_cast<void**>(&h->d_raw), bytes + 256));
        h->raw_capacity = bytes;
    }
    return MOT_OK;
}

static const int kBboxInit[8] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000, 0, 0};
int reset_bbox(mot_handle* h) {
    CK(cudaMemcpyAsync(h->d_bbox, kBboxInit, sizeof(kBboxInit), cudaMemcpyHostToDevice, h->stream));
    return MOT_OK;
}

// Start of a single-frame call: nothing is enqueued yet -- the small-frame graph clears its own counters, and frame_core
// calls reset_frame_state when the frame takes the general path.
void begin_frame(mot_handle* h) {
    h->prof.launches = 0;
    h->have_result = false;
    h->tim_stages = true;
}

int reset_frame_state(mot_handle* h) {
    static const int bbox_init[8] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000, 0, 0};
    CK(cudaMemsetAsync(h->d_counts, 0, CNT_N * sizeof(int), h->stream));
    CK(cudaMemcpyAsync(h->d_bbox, bbox_init, sizeof(bbox_init), cudaMemcpyHostToDevice, h->stream));
    h->prof.launches = 0;
    h->have_result = false;
    return MOT_OK;
}

// ------------------------------------------------------------------------------------------------------------
// clustering core.  cloud: device pointer to the (compacted) cloud.  m < 0: M is taken from d_counts[CNT_M]
// (written by k_rs_compact) and the bbox was produced by k_rs_count; otherwise bbox is computed here.
// ------------------------------------------------------------------------------------------------------------
template <typename KT>
int cluster_sorted_part(mot_handle* h, const float4* cloud, int M, const GridCodec& g, int total_bits, int n_frames) {
    cudaStream_t st = h->stream;
    KT* keys[2] = {reinterpret_cast<KT*>(h->d_keys[0]), reinterpret_cast<KT*>(h->d_keys[1])};
    h->uf_mode_now = h->uf_mode == 3 ? (M >= h->uf_auto_points ? 2 : 1) : h->uf_mode;
    if (h->keys_hist_fused && h->rws.mode == 0) {
        const RsPlan pl = rs_plan(M, total_bits, h->rws);
        LAUNCH(KID_KEYS, k_cell_keys_hist<KT><<<pl.ck.grid, RS_THREADS, ((size_t)1 << pl.bits0) * sizeof(unsigned), st>>>(cloud, M, pl.ck.chunk, g, h->d_frame_offsets,
                                                                                                                       keys[0], pl.bits0, h->rws.hist, h->d_counts + CNT_FLAGS));
        h->rws.first_hist_done = true;
    } else {
        LAUNCH(KID_KEYS, k_cell_keys<KT><<<(M + 255) / 256, 256, 0, st>>>(cloud, M, g, h->d_frame_offsets, keys[0], h->d_counts + CNT_FLAGS));
    }
    const int sb = radix_sort_pairs<KT>(st, keys, h->d_vals, M, total_bits, true, h->rws, h->prof, KID_SORT_HIST);
    h->rws.first_hist_done = false;
    const KT* skeys = keys[sb];
    const uint32_t* svals = h->d_vals[sb];
    h->sorted_buf = sb;

    // coarse-cell hash: capacity for the worst case, actual size chosen on the device (k_hash_clear)
    long long coarse_cap = (long long)g.ncx * g.ncy * g.ncz * n_frames;
    if (coarse_cap > M || coarse_cap <= 0) coarse_cap = M;
    int hb_max = ceil_log2(4 * coarse_cap);
    if (hb_max < 4) hb_max = 4;
    while (((size_t)1 << hb_max) > h->hash_capacity) --hb_max;  // capacity holds >= 2 slots per point: load <= 0.5 always
    const unsigned hmask = 0;  // both are read from d_counts[CNT_HB] inside the kernels
    const int hshift = 0;
    const Chunking ck = make_chunking(M, CELLW_TILE, CELLW_MAX_GRID);
    LAUNCH(KID_CELLS_COUNT, k_cells_count<KT><<<ck.grid, CELL_THREADS, 0, st>>>(skeys, M, ck.chunk, h->d_blk));
    LAUNCH(KID_HASH_CLEAR, k_hash_clear<KT><<<h->num_sms * 2, 256, 0, st>>>(h->d_blk, ck.grid, hb_max, reinterpret_cast<KT*>(h->d_hkeys), h->d_counts,
                                                                             h->d_blk + 2 * ck.grid));
    LAUNCH(KID_CELLS_WRITE, k_cells_write<KT><<<ck.grid, CELL_THREADS, 0, st>>>(skeys, svals, cloud, h->d_spts, M, ck.chunk, h->d_blk, h->d_fc_start,
                                                                               h->d_cc_first, h->d_parent, h->d_csize, h->d_cmin, h->d_crank,
                                                                               h->d_counts, reinterpret_cast<KT*>(h->d_ckey),
                                                                               h->uf_mode_now == 2 ? h->d_fcode : nullptr));
    {
        int hgrid = (M + 255) / 256;
        if (hgrid > h->num_sms * 8) hgrid = h->num_sms * 8;
        LAUNCH(KID_HASH_BUILD, k_hash_build<KT><<<hgrid, 256, 0, st>>>(reinterpret_cast<const KT*>(h->d_ckey), h->d_counts,
                                                                        reinterpret_cast<KT*>(h->d_hkeys), h->d_hvals));
    }
    CK(cudaEventRecord(h->ev[2], st));

    const float r2 = (float)((double)h->tol * (double)h->tol);  // KdTreeFLANN::radiusSearch: (float)(radius*radius)
    int uf_grid = (M + UF_WARPS - 1) / UF_WARPS;
    if (uf_grid > h->num_sms * 8) uf_grid = h->num_sms * 8;
    int flat_grid = (M + 255) / 256;
    if (flat_grid > h->num_sms * 8) flat_grid = h->num_sms * 8;
    if (h->uf_mode_now == 2) {
        const cudaStream_t main_st = st;
        const bool side = h->uf_stream && !h->prof.on;
        if (side) {
            CK(cudaEventRecord(h->ev_uf[0], main_st));
            CK(cudaStreamWaitEvent(h->uf_stream, h->ev_uf[0], 0));
            st = h->uf_stream;
        }
        if (h->l2_persist_mb > 0) {  // keep the randomly accessed table of the walk resident in L2 across the five row sweeps
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = h->l2_persist_what == 0 ? (void*)h->d_parent : (void*)h->d_crec;
            av.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)M * (h->l2_persist_what == 0 ? 1 : 2), (size_t)h->l2_persist_mb << 20);
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        int lgrid = (M + CLOC_THREADS - 1) / CLOC_THREADS;
        if (lgrid > h->num_sms * 7) lgrid = h->num_sms * 7;
        int* dlist = h->cell_dense ? h->d_dense_list : nullptr;
        LAUNCH(KID_CELL_LOCAL, k_cell_local<<<lgrid, CLOC_THREADS, 0, st>>>(h->d_spts, h->d_fc_start, h->d_cc_first, h->d_fcode, h->d_counts, r2, h->uf_light,
                                                                           h->d_crec, h->uf_xmode == 0 ? h->d_cbox : nullptr, h->d_fbox, h->d_parent, h->d_heavy1, h->d_heavy2,
                                                                           h->heavy_cap, dlist, h->dense_cap));
        if (dlist) {
            int dgrid = (M + CLOC_DENSE_POINTS - 1) / CLOC_DENSE_POINTS;
            const int per_sm = 1536 / h->cld_threads;
            if (dgrid > h->num_sms * per_sm) dgrid = h->num_sms * per_sm;
            LAUNCH(KID_CELL_LOCAL_DENSE, k_cell_local_dense<<<dgrid, h->cld_threads, 0, st>>>(h->d_spts, h->d_fc_start, h->d_cc_first, h->d_fcode, h->d_counts, r2,
                                                                                         h->uf_light, h->d_crec, h->uf_xmode == 0 ? h->d_cbox : nullptr, h->d_fbox, h->d_parent,
                                                                                         h->d_heavy1, h->d_heavy2, h->heavy_cap, h->d_dense_list,
                                                                                         h->dense_cap));
        }
        if (h->uf_xmode == 2) {
            int fgrid = (M + UFF_THREADS - 1) / UFF_THREADS;
            if (fgrid > h->num_sms * h->uf_fused_blocks) fgrid = h->num_sms * h->uf_fused_blocks;
            LAUNCH(KID_UF_FUSED, k_uf_fused<KT><<<fgrid, UFF_THREADS, 0, st>>>(reinterpret_cast<const KT*>(h->d_ckey), h->d_crec, h->d_fbox, h->d_spts,
                                                                              reinterpret_cast<const KT*>(h->d_hkeys), h->d_hvals, h->d_counts, h->d_parent,
                                                                              g, r2, h->uf_light, h->d_heavy1, h->d_heavy2, h->heavy_cap, h->uf_row_inner, h->uf_tile_batches));
        } else if (h->uf_xmode == 1) {
            int sgrid = (M + UFS_THREADS - 1) / UFS_THREADS;
            if (sgrid > h->num_sms * 8) sgrid = h->num_sms * 8;
            int wgrid = (M + UFW_THREADS - 1) / UFW_THREADS;
            if (wgrid > h->num_sms * h->uf_walk_blocks) wgrid = h->num_sms * h->uf_walk_blocks;
            for (int ph = 0; ph < h->uf_n_phases; ++ph) {
                LAUNCH(KID_UF_SURV, k_uf_survivors<KT><<<sgrid, UFS_THREADS, 0, st>>>(reinterpret_cast<const KT*>(h->d_ckey), h->d_crec,
                                                                                     reinterpret_cast<const KT*>(h->d_hkeys), h->d_hvals, h->d_counts,
                                                                                     h->d_parent, g, h->uf_phases[ph], h->d_tasks, h->task_cap,
                                                                                     h->d_counts + CNT_TASKS0 + ph));
                LAUNCH(KID_UF_WALK, k_uf_walk<<<wgrid, UFW_THREADS, 0, st>>>(h->d_tasks, h->d_counts + CNT_TASKS0 + ph, h->task_cap,
                                                                            h->d_counts + CNT_TICKET0 + ph, h->d_crec, h->d_fbox, h->d_spts, h->d_parent, r2,
                                                                            h->uf_light, h->d_heavy1, h->d_heavy2, h->heavy_cap, h->d_counts));
            }
        } else {
        int xgrid = (M + UFX_THREADS - 1) / UFX_THREADS;
        if (xgrid > h->num_sms * h->uf_cross_blocks) xgrid = h->num_sms * h->uf_cross_blocks;
        const int n_split = h->uf_split ? 2 : 1;
        for (int part = 0; part < n_split; ++part) {
            const int rb = n_split == 1 ? 0 : (part == 0 ? 0 : 3), re = n_split == 1 ? 5 : (part == 0 ? 3 : 5);
            LAUNCH(KID_UF_CROSS, k_uf_cross<KT><<<xgrid, UFX_THREADS, 0, st>>>(reinterpret_cast<const KT*>(h->d_ckey), h->d_crec, h->d_cbox, h->d_fbox,
                                                                              h->d_spts, reinterpret_cast<const KT*>(h->d_hkeys), h->d_hvals, h->d_counts,
                                                                              h->d_parent, g, r2, h->uf_light, h->d_heavy1, h->d_heavy2, h->heavy_cap, rb, re));
        }
        }
        const int hgrid = h->num_sms * 8;
        LAUNCH(KID_UF_HEAVY1, k_uf_heavy<<<hgrid, UFH_THREADS, 0, st>>>(h->d_spts, h->d_fbox, h->d_heavy1, h->d_counts + CNT_HEAVY1, h->heavy_cap, h->d_parent, r2));
        LAUNCH(KID_FLATTEN_IF, k_uf_flatten_if<<<flat_grid, 256, 0, st>>>(h->d_parent, h->d_counts, h->d_counts + CNT_HEAVY2));
        LAUNCH(KID_UF_HEAVY2, k_uf_heavy<<<hgrid, UFH_THREADS, 0, st>>>(h->d_spts, h->d_fbox, h->d_heavy2, h->d_counts + CNT_HEAVY2, h->heavy_cap, h->d_parent, r2));
        LAUNCH(KID_FLATTEN2, k_uf_flatten<false><<<flat_grid, 256, 0, st>>>(h->d_parent, h->d_root, h->d_counts));
        if (side) {
            CK(cudaEventRecord(h->ev_uf[1], st));
            st = main_st;
            CK(cudaStreamWaitEvent(st, h->ev_uf[1], 0));
        }
    } else if (h->uf_mode_now == 1) {
        int rgrid = (M + 15) / 16;
        if (rgrid > h->num_sms * 32) rgrid = h->num_sms * 32;
        LAUNCH(KID_COARSE_REC, k_coarse_records<KT><<<rgrid, 256, 0, st>>>(skeys, h->d_fc_start, h->d_cc_first, h->d_counts,
                                                                           reinterpret_cast<const KT*>(h->d_hkeys), h->d_hvals, hmask, hshift, g,
                                                                           h->d_crec, h->d_nbr));
        int cgrid = (M + UFC_WARPS - 1) / UFC_WARPS;
        if (cgrid > h->num_sms * h->uf_blocks_per_sm) cgrid = h->num_sms * h->uf_blocks_per_sm;
        const size_t usm = UFC_WARPS * sizeof(UfcWarpSmem);
        const cudaStream_t main_st = st;
        const bool side = h->uf_stream && !h->prof.on;
        if (side) {
            CK(cudaEventRecord(h->ev_uf[0], main_st));
            CK(cudaStreamWaitEvent(h->uf_stream, h->ev_uf[0], 0));
            st = h->uf_stream;
        }
        if (h->uf_pair)
            LAUNCH(KID_UF_PAIR, k_uf_sparse2<<<cgrid, UFC_THREADS, UFC_WARPS * sizeof(UfpWarpSmem), st>>>(h->d_spts, h->d_crec, h->d_nbr, h->d_counts,
                                                                                                      h->d_parent, r2, h->uf_tma, h->d_dense_list,
                                                                                                      h->dense_cap));
        else
            LAUNCH(KID_UF_COARSE, k_uf_sparse<<<cgrid, UFC_THREADS, usm, st>>>(h->d_spts, h->d_fc_start, h->d_crec, h->d_nbr, h->d_counts, h->d_parent, r2,
                                                                               h->uf_tma, h->d_dense_list, h->dense_cap));
        int dgrid = cgrid < h->num_sms * 8 ? cgrid : h->num_sms * 8;
        LAUNCH(KID_UF_DENSE1, k_uf_dense<1><<<dgrid, UFC_THREADS, 0, st>>>(h->d_spts, h->d_fc_start, h->d_crec, h->d_nbr, h->d_counts, h->d_parent, r2,
                                                                            h->d_dense_list, h->dense_cap));
        LAUNCH(KID_FLATTEN1, k_uf_flatten<true><<<flat_grid, 256, 0, st>>>(h->d_parent, h->d_root, h->d_counts));
        LAUNCH(KID_UF_DENSE, k_uf_dense<2><<<dgrid, UFC_THREADS, 0, st>>>(h->d_spts, h->d_fc_start, h->d_crec, h->d_nbr, h->d_counts, h->d_parent, r2,
                                                                           h->d_dense_list, h->dense_cap));
        LAUNCH(KID_FLATTEN2, k_uf_flatten<false><<<flat_grid, 256, 0, st>>>(h->d_parent, h->d_root, h->d_counts));
        if (side) {
            CK(cudaEventRecord(h->ev_uf[1], st));
            st = main_st;
            CK(cudaStreamWaitEvent(st, h->ev_uf[1], 0));
        }
    } else {
        LAUNCH(KID_UF1, k_uf_pairs<KT, 1><<<uf_grid, UF_THREADS, 0, st>>>(skeys, h->d_spts, h->d_fc_start, h->d_cc_first,
                                                                          reinterpret_cast<const KT*>(h->d_hkeys), h->d_hvals, hmask, hshift,
                                                                          h->d_counts, h->d_parent, g, r2));
        LAUNCH(KID_FLATTEN1, k_uf_flatten<true><<<flat_grid, 256, 0, st>>>(h->d_parent, h->d_root, h->d_counts));
        LAUNCH(KID_UF2, k_uf_pairs<KT, 2><<<uf_grid, UF_THREADS, 0, st>>>(skeys, h->d_spts, h->d_fc_start, h->d_cc_first,
                                                                          reinterpret_cast<const KT*>(h->d_hkeys), h->d_hvals, hmask, hshift,
                                                                          h->d_counts, h->d_parent, g, r2));
        LAUNCH(KID_FLATTEN2, k_uf_flatten<false><<<flat_grid, 256, 0, st>>>(h->d_parent, h->d_root, h->d_counts));
    }
    CK(cudaEventRecord(h->ev[3], st));
    CK(cudaGetLastError());
    return MOT_OK;
}

int cluster_core(mot_handle* h, const float4* cloud, int m_known, int n_frames, bool with_centroids, double stamp) {
    cudaStream_t st = h->stream;
    // Speculation: a sensor's extent hardly changes from call to call, so the grid of the previous call (padded, see below) is
    // tried first -- no bounding-box kernel, no host round trip before the keys.  k_cell_keys verifies every point (flag 32).
    const bool spec = h->plan_spec && h->plan_valid && m_known > 0 && h->plan_g.n_frames == n_frames && h->plan_tol == h->tol;
    GridCodec g{};
    int total_bits = 0;
    int M = m_known;
    if (!spec) {
        if (m_known > 0) {
            int grid = (m_known + 255) / 256;
            if (grid > h->num_sms * 8) grid = h->num_sms * 8;
            LAUNCH(KID_BBOX, k_bbox<<<grid, 256, 0, st>>>(cloud, m_known, h->d_bbox));
        }
        CK(cudaEventRecord(h->ev[1], st));
        // ---- S1 ----
        CK(cudaMemcpyAsync(h->h_pinned, h->d_bbox, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h->h_pinned + 8, h->d_counts, CNT_N * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(mot_sync(h));
        M = m_known >= 0 ? m_known : h->h_pinned[8 + CNT_M];
    } else {
        CK(cudaEventRecord(h->ev[1], st));
    }
    h->res_cloud = cloud;
    h->res_M = M;
    h->res_K = 0;
    h->res_total = 0;
    h->res_frames = n_frames;
    h->res_centroids = false;
    h->res_fine = h->res_coarse = h->res_key_bits = 0;
    if (!spec && h->h_pinned[6] != 0) return fail(h, MOT_ERR_NONFINITE, "cloud contains NaN/Inf coordinates");
    if (M == 0) {
        for (int i = 1; i < 6; ++i) CK(cudaEventRecord(h->ev[i], st));
        CK(cudaMemsetAsync(h->d_cl_offsets, 0, sizeof(int), st));
        if (n_frames > 1) CK(cudaMemsetAsync(h->d_frame_cl_offsets, 0, (n_frames + 1) * sizeof(int), st));
        h->have_result = true;
        return MOT_OK;
    }
    const int frame_bits = n_frames > 1 ? ceil_log2(n_frames) : 0;
    if (spec) {
        g = h->plan_g;
        total_bits = h->plan_bits;
    } else {
        float mn[3], mx[3];
        for (int d = 0; d < 3; ++d) {
            mn[d] = ordered_to_float_bits(h->h_pinned[d]);
            mx[d] = ordered_to_float_bits(h->h_pinned[3 + d]);
        }
        const double hcell = (double)h->tol * (1.0 + 1.0 / 1024.0);  // coarse edge: tol * (1 + 2^-10)
        const double e = hcell * 0.5;                               // fine (clique) edge
        g.inv_e = 1.0 / e;
        g.minx = mn[0]; g.miny = mn[1]; g.minz = mn[2];
        long long nf[3];
        for (int d = 0; d < 3; ++d) {
            const double span = ((double)mx[d] - (double)mn[d]) * g.inv_e;
            if (!(span < 1.0e9)) return fail(h, MOT_ERR_INVALID, "cloud extent / cluster_tolerance too large for the voxel grid");
            nf[d] = (long long)std::floor(span) + 1;
        }
        g.nfx = (int)nf[0]; g.nfy = (int)nf[1]; g.nfz = (int)nf[2];
        g.ncx = (g.nfx + 1) / 2; g.ncy = (g.nfy + 1) / 2; g.ncz = (g.nfz + 1) / 2;
        g.bx = ceil_log2(g.ncx); g.by = ceil_log2(g.ncy); g.bz = ceil_log2(g.ncz);
        g.n_frames = n_frames;
        total_bits = 3 + g.bx + g.by + g.bz + frame_bits;
        if (total_bits > 63) return fail(h, MOT_ERR_INVALID, "voxel key does not fit in 64 bits");
        // The plan for the next call: the same key layout, every axis widened to the full range of its bits (and by the bits
        // that are free up to the next radix pass), the origin moved so that the slack lies on both sides of this call's box.
        {
            GridCodec pg = g;
            int bits[3] = {g.bx, g.by, g.bz};
            const int passes = (total_bits + RS_MAX_BITS - 1) / RS_MAX_BITS;
            int spare = std::min(passes * RS_MAX_BITS, total_bits <= 32 ? 32 : 63) - total_bits;
            const int nc00[3] = {g.ncx, g.ncy, g.ncz};
            for (int d = 0; d < 3 && spare > 0; ++d)  // an axis that fills its bits exactly has no slack: it gets a free bit
                if ((1 << bits[d]) == nc00[d]) {
                    ++bits[d];
                    --spare;
                }
            const int nc0[3] = {g.ncx, g.ncy, g.ncz};
            double pmin[3] = {g.minx, g.miny, g.minz};
            int ncp[3];
            for (int d = 0; d < 3; ++d) {
                ncp[d] = 1 << bits[d];
                pmin[d] -= (double)((ncp[d] - nc0[d]) / 2) * hcell;
            }
            pg.minx = pmin[0]; pg.miny = pmin[1]; pg.minz = pmin[2];
            pg.bx = bits[0]; pg.by = bits[1]; pg.bz = bits[2];
            pg.ncx = ncp[0]; pg.ncy = ncp[1]; pg.ncz = ncp[2];
            pg.nfx = 2 * ncp[0]; pg.nfy = 2 * ncp[1]; pg.nfz = 2 * ncp[2];
            h->plan_g = pg;
            h->plan_bits = 3 + bits[0] + bits[1] + bits[2] + frame_bits;
            h->plan_tol = h->tol;
            h->plan_valid = h->plan_bits <= 63;
        }
    }
    int rc = total_bits <= 32 ? cluster_sorted_part<uint32_t>(h, cloud, M, g, total_bits, n_frames)
                              : cluster_sorted_part<uint64_t>(h, cloud, M, g, total_bits, n_frames);
    if (rc != MOT_OK) return rc;

    // ---- K6: sizes, filter ----
    ClusterKeyCodec kc{};
    long long max_frame_len = M;  // single frame
    kc.size_bits = ceil_log2((long long)std::min<long long>(h->max_size, M) + 1);
    if (kc.size_bits < 1) kc.size_bits = 1;
    kc.size_cap = (unsigned)((1ull << kc.size_bits) - 1);
    kc.idx_bits = ceil_log2(max_frame_len + 1);
    if (kc.idx_bits < 1) kc.idx_bits = 1;
    const int ckey_bits = kc.size_bits + kc.idx_bits + frame_bits;
    if (ckey_bits > 64) return fail(h, MOT_ERR_INVALID, "cluster ordering key does not fit in 64 bits (max_cluster_size x points x frames too large)");
    int cgrid = (M + 255) / 256;
    if (cgrid > h->num_sms * 8) cgrid = h->num_sms * 8;
    LAUNCH(KID_COMP_ACC, k_comp_accumulate<<<cgrid, 256, 0, st>>>(h->d_fc_start, h->d_root, h->d_csize, h->d_cmin, h->d_counts));
    LAUNCH(KID_KEPT_LIST, k_kept_list<<<cgrid, 256, 0, st>>>(h->d_root, h->d_csize, h->d_cmin, h->d_frame_offsets, n_frames, h->min_size,
                                                            h->max_size, kc, h->d_ckeys[0], h->d_croots[0], h->d_counts));
    // ---- S2 ----
    CK(cudaMemcpyAsync(h->h_pinned + 8, h->d_counts, CNT_N * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(mot_sync(h));
    const int K = h->h_pinned[8 + CNT_K];
    const int total = h->h_pinned[8 + CNT_TOTAL];
    h->res_K = K;
    h->res_total = total;
    std::memcpy(h->res_counters, h->h_pinned + 8, sizeof(h->res_counters));
    if (h->h_pinned[8 + CNT_FLAGS] & 64) return fail(h, MOT_ERR_NONFINITE, "cloud contains NaN/Inf coordinates");
    if (h->h_pinned[8 + CNT_FLAGS] & 32) {
        if (!spec) return fail(h, MOT_ERR_CUDA, "internal: a point fell outside the grid planned from its own bounding box");
        // the cloud outgrew the speculative grid: plan from this call's bounding box and run again (the work so far is void)
        h->plan_valid = false;
        ++h->plan_misses;
        const int launches_so_far = h->prof.launches;
        int rc2 = reset_frame_state(h);
        if (rc2 != MOT_OK) return rc2;
        h->prof.launches = launches_so_far;
        return cluster_core(h, cloud, m_known, n_frames, with_centroids, stamp);
    }
    if (spec) ++h->plan_hits;
    if (h->h_pinned[8 + CNT_FLAGS] & 1) return fail(h, MOT_ERR_CAPACITY, "dense-task list overflow (internal capacity)");
    if (h->h_pinned[8 + CNT_FLAGS] & 16) return fail(h, MOT_ERR_CUDA, "internal self check failed (MOT_CHECKS build): an index left its bounds in cell_uf.cuh");
    if (h->h_pinned[8 + CNT_FLAGS] & 8) return fail(h, MOT_ERR_CUDA, "TMA staging of a cell tile timed out (mbarrier never completed)");
    if (h->h_pinned[8 + CNT_FLAGS] & 4) return fail(h, MOT_ERR_CAPACITY, "cell-pair task list overflow (internal capacity)");
    if (h->h_pinned[8 + CNT_FLAGS] & 2) return fail(h, MOT_ERR_CUDA, "radix sort look-back exceeded its spin limit");
    h->res_fine = h->h_pinned[8 + CNT_FINE];
    h->res_coarse = h->h_pinned[8 + CNT_COARSE];
    h->res_key_bits = total_bits;

    const uint64_t* sorted_ckeys = nullptr;
    if (K <= CL_SMALL_MAX) {
        LAUNCH(KID_CL_SMALL, k_clusters_small<<<1, CL_SMALL_THREADS, CL_SMALL_SMEM, st>>>(h->d_ckeys[0], h->d_croots[0], K, kc, h->d_ckeys[1],
                                                                                         h->d_crank, h->d_cl_offsets));
        sorted_ckeys = h->d_ckeys[1];
    } else {
        const int cb = radix_sort_pairs<uint64_t>(st, h->d_ckeys, h->d_croots, K, ckey_bits, false, h->rws, h->prof, KID_CLSORT_HIST);
        const Chunking ck = make_chunking(K, CLF_THREADS, CLF_MAX_GRID);
        LAUNCH(KID_CL_COUNT, k_clusters_count<<<ck.grid, CLF_THREADS, 0, st>>>(h->d_ckeys[cb], K, ck.chunk, kc, h->d_blk));
        LAUNCH(KID_CL_FINALIZE, k_clusters_finalize<<<ck.grid, CLF_THREADS, 0, st>>>(h->d_ckeys[cb], h->d_croots[cb], K, ck.chunk, kc, h->d_blk,
                                                                                    h->d_crank, h->d_cl_offsets));
        sorted_ckeys = h->d_ckeys[cb];
    }
    if (n_frames > 1) {
        LAUNCH(KID_FRAME_CL_OFF, k_frame_cluster_offsets<<<(n_frames + 1 + 127) / 128, 128, 0, st>>>(sorted_ckeys, K, kc, n_frames,
                                                                                                    h->d_frame_cl_offsets));
    }
    // ---- CSR emission: stable partition of 0..M-1 by cluster rank ----
    uint32_t* pk[2] = {reinterpret_cast<uint32_t*>(h->d_keys[0]), reinterpret_cast<uint32_t*>(h->d_keys[1])};
    LAUNCH(KID_POINT_RANK, k_point_rank<<<(M + 255) / 256, 256, 0, st>>>(h->d_spts, h->d_vals[h->sorted_buf], h->d_root, h->d_crank, h->d_cmin, M, K, pk[0],
                                                                        h->d_labels));
    if (h->csr_compact && (long long)total * 10 < (long long)M * 6) {  // pays off when the size filter drops >= 40 % of the points
        // only the points of kept clusters enter the sort: (rank, index) pairs in index order, then a stable sort by rank
        uint32_t* ck[2] = {pk[1], pk[0]};
        uint32_t* cv[2] = {h->d_vals[1], h->d_vals[0]};
        if (total > 0) {
            const int items = h->csr_compact == 8 ? 8 : 16;  // MOT_CSR_COMPACT=8: 2048-key tiles, else 4096
            const int tiles = (M + C1P_THREADS * items - 1) / (C1P_THREADS * items);
            CK(cudaMemsetAsync(h->d_c1p_status, 0, ((size_t)tiles + 1) * sizeof(unsigned), st));
            if (items == 8)
                LAUNCH(KID_CSR_COMPACT, k_compact_keys_onepass<8><<<tiles, C1P_THREADS, 0, st>>>(pk[0], M, (uint32_t)K, ck[0], cv[0], h->d_c1p_status, tiles, h->d_counts + CNT_FLAGS));
            else
                LAUNCH(KID_CSR_COMPACT, k_compact_keys_onepass<16><<<tiles, C1P_THREADS, 0, st>>>(pk[0], M, (uint32_t)K, ck[0], cv[0], h->d_c1p_status, tiles, h->d_counts + CNT_FLAGS));
            const int cb = radix_sort_pairs<uint32_t>(st, ck, cv, total, ceil_log2((long long)K + 1), false, h->rws, h->prof, KID_PART_HIST);
            h->res_idx_buf = cb == 0 ? 1 : 0;
        } else {
            h->res_idx_buf = 1;
        }
    } else {
        h->res_idx_buf = radix_sort_pairs<uint32_t>(st, pk, h->d_vals, M, ceil_log2((long long)K + 1), true, h->rws, h->prof, KID_PART_HIST);
    }
    CK(cudaEventRecord(h->ev[4], st));

    // ---- K7 / K8 (on global point indices; batch indices are made frame-local afterwards) ----
    if (K > 0) {
        int slabs = 1;
        const bool cent = with_centroids;
        if (cent && K < 2000) {
            slabs = (h->num_sms * 16 + K - 1) / K;
            if (slabs > 64) slabs = 64;
            if (slabs < 1) slabs = 1;
        }
        rc = ensure_tables(h, (size_t)K, cent ? (size_t)K * slabs : 0);
        if (rc != MOT_OK) return rc;
        int sgrid = K < h->num_sms * 16 ? K : h->num_sms * 16;
        LAUNCH(KID_STATS_INIT, k_stats_init<<<(K + 255) / 256, 256, 0, st>>>(h->d_statacc, K));
        const int sg = stat_groups(total, h->num_sms);
        const int per_block = 32 * sg * (STAT_THREADS / 32);
        LAUNCH(KID_STATS, k_stats_accumulate<<<(total + per_block - 1) / per_block, STAT_THREADS, 0, st>>>(cloud, h->d_cl_offsets, h->d_vals[h->res_idx_buf],
                                                                                                         K, total, h->d_statacc, sg));
        LAUNCH(KID_STATS_FIN, k_stats_finalize<<<(K + 255) / 256, 256, 0, st>>>(h->d_statacc, h->d_cl_offsets, K, h->d_stats));
        if (cent) {
            LAUNCH(KID_FARTHEST_PAIR, k_farthest_pair<<<K * slabs, FP_THREADS, 0, st>>>(cloud, h->d_cl_offsets, h->d_vals[h->res_idx_buf], K, slabs,
                                                                                       h->d_cands));
            LAUNCH(KID_CIRCUMCENTRE, k_circumcentre<<<sgrid, CC_THREADS, 0, st>>>(cloud, h->d_cl_offsets, h->d_vals[h->res_idx_buf], K, slabs,
                                                                                 h->d_cands, (float)stamp, h->d_centroids,
                                                                                 n_frames > 1 && h->have_frame_stamps ? h->d_frame_stamps : nullptr,
                                                                                 h->d_frame_cl_offsets, n_frames));
            h->res_centroids = true;
        }
    }
    if (n_frames > 1 && total > 0) {
        LAUNCH(KID_LOCALIZE, k_localize_indices<<<(total + 255) / 256, 256, 0, st>>>(h->d_vals[h->res_idx_buf], total, h->d_frame_offsets, n_frames));
    }
    CK(cudaEventRecord(h->ev[5], st));
    CK(cudaGetLastError());
    h->have_result = true;
    return MOT_OK;
}

// accumulate the per-kernel event pairs of the call that just finished (stream already synchronised)
void fold_profile(mot_handle* h) {
    if (h->prof.on) {
        for (auto& r : h->prof.recs) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { h->prof_ms[r.id] += ms; h->prof_n[r.id] += 1; }
        }
        cudaGetLastError();
    }
    h->prof.recs.clear();
    h->prof.used = 0;
}

int finish_timings(mot_handle* h) {
    CK(mot_sync(h));
    float t01 = 0, t12 = 0, t23 = 0, t34 = 0, t45 = 0, t05 = 0;
    if (!h->tim_stages) {  // small-frame path: one graph, only the total is measured
        cudaEventElapsedTime(&t05, h->ev[0], h->ev[5]);
        cudaGetLastError();
        h->tim = mot_timings{};
        h->tim.total_ms = t05;
        fold_profile(h);
        return MOT_OK;
    }
    cudaEventElapsedTime(&t01, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&t12, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&t23, h->ev[2], h->ev[3]);
    cudaEventElapsedTime(&t34, h->ev[3], h->ev[4]);
    cudaEventElapsedTime(&t45, h->ev[4], h->ev[5]);
    cudaEventElapsedTime(&t05, h->ev[0], h->ev[5]);
    cudaGetLastError();  // a frame that ended early (M == 0 / error) may not have recorded every event
    h->tim.remove_static_ms = t01;
    h->tim.grid_build_ms = t12;
    h->tim.union_find_ms = t23;
    h->tim.cluster_table_ms = t34;
    h->tim.reduce_ms = t45;
    h->tim.total_ms = t05;
    fold_profile(h);
    return MOT_OK;
}

constexpr size_t STAGE_IN = 2u << 20, STAGE_OUT = 2u << 20, STAGE_CHUNK = 256u << 10;

// true for ordinary (pageable, unregistered) host memory; pinned, managed and device pointers take the direct copies
bool is_pageable_host(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// caller's cloud -> device (stream ordered).  A small pageable source is copied chunk by chunk into the pinned staging, every chunk's DMA
// running while the next one is copied; the staging is free again when the call has synchronised (every host entry does).
int upload_cloud(mot_handle* h, void* d_dst, const void* src, size_t bytes) {
    if (!bytes) return MOT_OK;
    if (h->h_stage && bytes <= STAGE_IN && is_pageable_host(src)) {
        for (size_t o = 0; o < bytes; o += STAGE_CHUNK) {
            const size_t c = std::min(STAGE_CHUNK, bytes - o);
            std::memcpy(h->h_stage + o, static_cast<const unsigned char*>(src) + o, c);
            CK(cudaMemcpyAsync(static_cast<unsigned char*>(d_dst) + o, h->h_stage + o, c, cudaMemcpyHostToDevice, h->stream));
        }
        return MOT_OK;
    }
    CK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyDefault, h->stream));
    return MOT_OK;
}

int fetch_result(mot_handle* h, float* kept, size_t kept_cap, int32_t* offs, size_t offs_cap, int32_t* idx, size_t idx_cap,
                 mot_cluster_stat* stats, float* cent, size_t table_cap) {
    if (!h->have_result) return fail(h, MOT_ERR_STATE, "no clustering result on this handle");
    cudaStream_t st = h->stream;
    if (kept && kept_cap < (size_t)h->res_M) return fail(h, MOT_ERR_CAPACITY, "kept cloud buffer too small");
    if (offs && offs_cap < (size_t)h->res_K + 1) return fail(h, MOT_ERR_CAPACITY, "cluster_offsets buffer too small");
    if (idx && idx_cap < (size_t)h->res_total) return fail(h, MOT_ERR_CAPACITY, "point_indices buffer too small");
    if (stats && h->res_K && table_cap < (size_t)h->res_K) return fail(h, MOT_ERR_CAPACITY, "stats buffer too small");
    if (cent && h->res_K) {
        if (table_cap < (size_t)h->res_K) return fail(h, MOT_ERR_CAPACITY, "centroid buffer too small");
        if (!h->res_centroids) return fail(h, MOT_ERR_STATE, "centroids were not computed for the last result");
    }
    struct Part { void* dst; const void* src; size_t bytes; };
    const Part parts[5] = {
        {kept, h->res_cloud, kept ? (size_t)h->res_M * 16 : 0},
        {offs, h->d_cl_offsets, offs ? ((size_t)h->res_K + 1) * 4 : 0},
        {idx, h->d_vals[h->res_idx_buf], idx ? (size_t)h->res_total * 4 : 0},
        {stats, h->d_stats, stats ? (size_t)h->res_K * sizeof(ClusterStat) : 0},
        {cent, h->d_centroids, cent ? (size_t)h->res_K * 16 : 0}};
    size_t total = 0;
    bool stage = h->h_stage != nullptr;
    for (const Part& p : parts)
        if (p.bytes) {
            total += (p.bytes + 63) & ~(size_t)63;
            stage = stage && total <= STAGE_OUT && is_pageable_host(p.dst);
        }
    if (stage && total) {  // all parts into the pinned staging (queued back to back), one wait, then plain copies to the caller's buffers
        unsigned char* out = h->h_stage + STAGE_IN;
        size_t o = 0;
        for (const Part& p : parts)
            if (p.bytes) {
                CK(cudaMemcpyAsync(out + o, p.src, p.bytes, cudaMemcpyDeviceToHost, st));
                o += (p.bytes + 63) & ~(size_t)63;
            }
        CK(mot_sync(h));
        o = 0;
        for (const Part& p : parts)
            if (p.bytes) {
                std::memcpy(p.dst, out + o, p.bytes);
                o += (p.bytes + 63) & ~(size_t)63;
            }
        return MOT_OK;
    }
    for (const Part& p : parts)
        if (p.bytes) CK(cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyDefault, st));
    CK(mot_sync(h));
    return MOT_OK;
}

int check_frame_args(mot_handle* h, const void* pts, size_t n) {
    if (!h) return MOT_ERR_INVALID;
    if (n > 0 && !pts) return fail(h, MOT_ERR_INVALID, "null point buffer");
    if (n > h->max_points) return fail(h, MOT_ERR_CAPACITY, "frame larger than the handle's max_points");
    if (n > 0x7ffffff0ull) return fail(h, MOT_ERR_CAPACITY, "frame too large");
    return MOT_OK;
}

// Small frames (frame_small.cuh): removeStatic + clustering + tables as five kernels launched back to back -- as ONE
// instantiated CUDA graph -- with a single host round trip at the end; every size in between (M, cells, K) stays on the device.
// The kernels of the small-frame path on `st`: arguments copied from their pinned staging, then the five kernels (the first
// clears the counters, the last stores them to pinned host memory).  Called directly (profiling / MOT_SMALL_GRAPH=0) or under stream capture (once per handle).
int fs_enqueue(mot_handle* h, cudaStream_t st, bool with_prof) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->fs_cluster);
    cfg.blockDim = dim3(FS_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = h->fs_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const FsArgs* ap = h->d_fs_args;
    CK(cudaMemcpyAsync(h->d_fs_args, h->h_fs_args, sizeof(FsArgs), cudaMemcpyHostToDevice, st));  // (k_fs_front clears the counters)
    if (with_prof) {
        LAUNCH(KID_FS_FRONT, CK(cudaLaunchKernelEx(&cfg, k_fs_front, ap)));
        LAUNCH(KID_FS_EDGES, k_fs_edges<<<h->num_sms * FS_PAIR_CTAS_PER_SM, FS_PAIR_THREADS, 0, st>>>(ap));
        cfg.dynamicSmemBytes = FS_TABLES_SMEM;
        LAUNCH(KID_FS_TABLES, CK(cudaLaunchKernelEx(&cfg, k_fs_tables, ap)));
        cfg.dynamicSmemBytes = 0;
        LAUNCH(KID_FS_FARTHEST, k_fs_farthest<<<h->num_sms * 2, FS_FP_THREADS, 0, st>>>(ap));
        LAUNCH(KID_FS_FINISH, k_fs_finish<<<h->num_sms * 2, FS_FIN_THREADS, 0, st>>>(ap));
    } else {
        CK(cudaLaunchKernelEx(&cfg, k_fs_front, ap));
        k_fs_edges<<<h->num_sms * FS_PAIR_CTAS_PER_SM, FS_PAIR_THREADS, 0, st>>>(ap);
        cfg.dynamicSmemBytes = FS_TABLES_SMEM;
        CK(cudaLaunchKernelEx(&cfg, k_fs_tables, ap));
        cfg.dynamicSmemBytes = 0;
        k_fs_farthest<<<h->num_sms * 2, FS_FP_THREADS, 0, st>>>(ap);
        k_fs_finish<<<h->num_sms * 2, FS_FIN_THREADS, 0, st>>>(ap);
    }
    return MOT_OK;  // (k_fs_finish stores the counters to the pinned host copy: no copy node)
}

int fs_build_graph(mot_handle* h) {
    cudaStream_t st = h->stream;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = fs_enqueue(h, st, false);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != MOT_OK || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        h->fs_use_graph = 0;  // plain launches from now on
        return rc != MOT_OK ? rc : MOT_OK;
    }
    const cudaError_t e2 = cudaGraphInstantiate(&h->fs_graph, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) {
        cudaGetLastError();
        h->fs_graph = nullptr;
        h->fs_use_graph = 0;
    }
    return MOT_OK;
}

// Returns MOT_OK (result ready), 1 (not taken), 2 (launched but handed back: the caller restarts the frame on the general
// path), < 0 on errors.
int frame_small(mot_handle* h, const float4* src, int n, bool do_rs, float4* rs_dst, bool with_centroids, double stamp) {
    if (h->fs_cluster == 0 || h->fs_max_points == 0 || n <= 0 || n > h->fs_max_points) return 1;
    if ((long long)n > (long long)h->fs_cluster * FS_THREADS * FS_MAX_ITEMS) return 1;
    if (h->fs_skip > 0) { --h->fs_skip; return 1; }
    const int wide_grid = h->num_sms * 8;
    int rc = ensure_tables(h, (size_t)std::min(n, FS_MAX_K), (size_t)FS_MAX_K + (size_t)wide_grid);
    if (rc != MOT_OK) return rc;
    cudaStream_t st = h->stream;
    FsArgs a{};
    a.src = src; a.n = n; a.do_rs = do_rs ? 1 : 0; a.mp = h->mp; a.bits = h->d_bits; a.kept = rs_dst;
    a.inv_e = 2.0 / ((double)h->tol * (1.0 + 1.0 / 1024.0));
    a.r2 = (float)((double)h->tol * (double)h->tol);
    a.min_size = h->min_size; a.max_size = h->max_size;
    a.with_centroids = with_centroids ? 1 : 0;
    a.stamp = (float)stamp;
    a.cell_cap = h->fs_cell_cap;
    a.fp_ctas = h->num_sms * 2;
    a.order_limit = with_centroids ? (1ull << 22) : (1ull << 25);
    a.log_t = std::max(4, ceil_log2(2 * (long long)n));
    a.T = 1 << a.log_t;  // <= hash_capacity (2^ceil_log2(2 max_points))
    a.hkeys = reinterpret_cast<unsigned long long*>(h->d_hkeys);
    a.hstart = h->d_hvals;
    a.pslot = reinterpret_cast<int*>(h->d_keys[0]);
    a.prank = reinterpret_cast<int*>(h->d_keys[1]);
    a.spts = h->d_spts;
    a.scell = reinterpret_cast<int*>(h->d_croots[1]);
    a.celllist = reinterpret_cast<int*>(h->d_crec);
    a.edges = reinterpret_cast<uint32_t*>(h->d_nbr);
    a.edge_cap = (int)std::min<size_t>(h->max_points * 16, 0x7fffffff);
    a.fbox = h->d_fbox;
    a.state = h->d_fs_state;
    a.parent = h->d_parent; a.root = h->d_root; a.csize = h->d_csize; a.cmin = h->d_cmin; a.crank = h->d_crank;
    a.ksize = reinterpret_cast<int*>(h->d_croots[0]);
    a.kmin = reinterpret_cast<int*>(h->d_ckeys[0]);
    a.kroot = reinterpret_cast<int*>(h->d_ckeys[1]);
    a.ssize = h->d_fc_start;
    a.cursor = h->d_cc_first;
    a.idx_tmp = h->d_vals[1];
    a.cands = h->d_cands;
    a.labels = h->d_labels; a.cl_offsets = h->d_cl_offsets; a.indices = h->d_vals[0];
    a.stats = h->d_stats; a.centroids = h->d_centroids; a.counts = h->d_counts;
    a.host_counts = h->h_pinned + 8;
    a.phase_ns = h->d_fs_phase_ns;
    *h->h_fs_args = a;  // the previous call on this handle has been synchronised: the staging is free
    if (h->fs_use_graph && !h->prof.on) {
        if (!h->fs_graph) {
            rc = fs_build_graph(h);
            if (rc != MOT_OK) return rc;
        }
    }
    if (h->fs_use_graph && !h->prof.on && h->fs_graph) {
        CK(cudaGraphLaunch(h->fs_graph, st));
        h->prof.launches += 5;
    } else {
        rc = fs_enqueue(h, st, true);
        if (rc != MOT_OK) return rc;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(h->ev[5], st));
    h->tim_stages = false;
    CK(mot_sync(h));
    const int flags = h->h_pinned[8 + CNT_FLAGS];
    if (flags & FS_FLAG_NONFINITE) return fail(h, MOT_ERR_NONFINITE, "cloud contains NaN/Inf coordinates");
    if (flags != 0) {  // FS_FLAG_FALLBACK: a crowded cell / too many clusters / coordinates beyond the hash key
        ++h->fs_misses;
        h->fs_skip = 15;
        return 2;
    }
    ++h->fs_hits;
    h->res_cloud = do_rs ? rs_dst : src;
    h->res_M = h->h_pinned[8 + CNT_M];
    h->res_K = h->h_pinned[8 + CNT_K];
    h->res_total = h->h_pinned[8 + CNT_TOTAL];
    h->res_frames = 1;
    h->res_centroids = with_centroids;
    h->res_idx_buf = 0;
    h->res_fine = h->h_pinned[8 + CNT_COARSE];
    h->res_coarse = 0;
    h->res_key_bits = 0;
    std::memcpy(h->res_counters, h->h_pinned + 8, sizeof(h->res_counters));
    h->have_result = true;
    return MOT_OK;
}

// One frame: removeStatic (optional, src -> rs_dst) + clustering.  Small frames first try the single-launch path.
int frame_core(mot_handle* h, const float4* src, int n, bool do_rs, float4* rs_dst, bool with_centroids, double stamp) {
    int rc = frame_small(h, src, n, do_rs, rs_dst, with_centroids, stamp);
    if (rc <= 0) return rc;
    {   // the general path starts from cleared counters (also after a small-frame launch that was handed back)
        const int launches = h->prof.launches;
        rc = reset_frame_state(h);
        if (rc != MOT_OK) return rc;
        h->prof.launches = launches;
        h->tim_stages = true;
    }
    if (do_rs && n > 0) {
        rc = enqueue_remove_static(h, src, n, rs_dst);
        if (rc != MOT_OK) return rc;
        return cluster_core(h, rs_dst, -1, 1, with_centroids, stamp);
    }
    return cluster_core(h, do_rs ? rs_dst : src, n, 1, with_centroids, stamp);
}

// compute centroids for an existing result that was produced without them
int compute_centroids_late(mot_handle* h, double stamp) {
    const int K = h->res_K;
    if (K == 0) { h->res_centroids = true; return MOT_OK; }
    int slabs = 1;
    if (K < 2000) {
        slabs = (h->num_sms * 16 + K - 1) / K;
        if (slabs > 64) slabs = 64;
    }
    int rc = ensure_tables(h, (size_t)K, (size_t)K * slabs);
    if (rc != MOT_OK) return rc;
    const int sgrid = K < h->num_sms * 16 ? K : h->num_sms * 16;
    LAUNCH(KID_FARTHEST_PAIR, k_farthest_pair<<<K * slabs, FP_THREADS, 0, h->stream>>>(h->res_cloud, h->d_cl_offsets, h->d_vals[h->res_idx_buf], K,
                                                                                      slabs, h->d_cands));
    LAUNCH(KID_CIRCUMCENTRE, k_circumcentre<<<sgrid, CC_THREADS, 0, h->stream>>>(h->res_cloud, h->d_cl_offsets, h->d_vals[h->res_idx_buf], K, slabs,
                                                                                h->d_cands, (float)stamp, h->d_centroids));
    CK(cudaGetLastError());
    h->res_centroids = true;
    return MOT_OK;
}

}  // namespace

// ================================================================================================================
extern "C" {

const char* mot_version(void) { return "mot_b200 0.1 (sm_100a)"; }

int mot_create(int device, size_t max_points, size_t max_tracks, mot_handle** out) {
    if (!out || max_points == 0 || max_points > ((size_t)1 << 27)) return MOT_ERR_INVALID;  // cell indices are packed into 27 bits / int row offsets
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return MOT_ERR_CUDA;
    mot_handle* h = new mot_handle();
    h->device = device;
    h->max_points = max_points;
    h->max_tracks = max_tracks;
    auto body = [&]() -> int {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        h->num_sms = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        if (const char* e = getenv("MOT_SYNC")) h->sync_mode = std::string(e) == "block" ? 2 : (std::string(e) == "yield" ? 1 : 0);
        if (const char* e = getenv("MOT_UF_BLOCKS")) h->uf_blocks_per_sm = std::min(8, std::max(1, atoi(e)));
        {
            const char* e = getenv("MOT_UF_PRIO");  // default on; MOT_UF_PRIO=0 keeps everything on one stream
            if (!e || atoi(e) != 0) {
                int least = 0, greatest = 0;
                CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
                CK(cudaStreamDestroy(h->stream));
                CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest));
                CK(cudaStreamCreateWithPriority(&h->uf_stream, cudaStreamNonBlocking, least));
                CK(cudaEventCreateWithFlags(&h->ev_uf[0], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&h->ev_uf[1], cudaEventDisableTiming));
            }
        }
        CK(cudaEventCreateWithFlags(&h->sync_ev, (h->sync_mode == 2 ? cudaEventBlockingSync : 0) | cudaEventDisableTiming));
        const size_t n = max_points;
        CK(dalloc(&h->d_in, n));
        CK(dalloc(&h->d_pts, n));
        CK(dalloc(&h->d_spts, n));
        for (int i = 0; i < 2; ++i) {
            CK(cudaMalloc(&h->d_keys[i], n * 8 + 256));
            CK(dalloc(&h->d_vals[i], n));
            CK(dalloc(&h->d_ckeys[i], n));
            CK(dalloc(&h->d_croots[i], n));
        }
        CK(dalloc(&h->d_fc_start, n + 1));
        CK(dalloc(&h->d_cc_first, n + 1));
        CK(dalloc(&h->d_parent, n));
        CK(dalloc(&h->d_root, n));
        CK(dalloc(&h->d_csize, n));
        CK(dalloc(&h->d_cmin, n));
        CK(dalloc(&h->d_crank, n));
        CK(dalloc(&h->d_labels, n));
        CK(dalloc(&h->d_cl_offsets, n + 1));
        CK(dalloc(&h->d_crec, n));
        h->dense_cap = (int)(n / 4 + 1024);  // a dense task's forward neighbourhood holds >= 64 points and a point lies in <= 14 of them
        CK(dalloc(&h->d_dense_list, (size_t)h->dense_cap));
        CK(dalloc(&h->d_nbr, n * 16));
        CK(cudaMalloc(&h->d_ckey, (n + 1) * 8 + 256));
        CK(dalloc(&h->d_fcode, n));
        if (const char* e = getenv("MOT_UF_XMODE")) h->uf_xmode = atoi(e);
        if (h->uf_xmode == 0) CK(dalloc(&h->d_cbox, 2 * n));  // the A/B variants' tables are only allocated when selected
        CK(dalloc(&h->d_fbox, 2 * n));
        h->heavy_cap = (int)(n / 2 + 4096);
        CK(dalloc(&h->d_heavy1, (size_t)h->heavy_cap));
        CK(dalloc(&h->d_heavy2, (size_t)h->heavy_cap));
        if (const char* e = getenv("MOT_L2_PERSIST")) h->l2_persist_mb = std::max(0, atoi(e));
        if (const char* e = getenv("MOT_L2_WHAT")) h->l2_persist_what = atoi(e);
        if (h->l2_persist_mb > 0) CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)h->l2_persist_mb << 20));
        if (const char* e = getenv("MOT_UF_TILE")) h->uf_tile_batches = std::max(0, atoi(e));
        if (const char* e = getenv("MOT_UF_ROWINNER")) h->uf_row_inner = atoi(e);
        if (const char* e = getenv("MOT_CLD_THREADS")) { const int v = atoi(e); if (v == 128 || v == 256 || v == 512) h->cld_threads = v; }
        if (const char* e = getenv("MOT_CELL_DENSE")) h->cell_dense = atoi(e);
        if (const char* e = getenv("MOT_UF_AUTO")) h->uf_auto_points = std::max(0, atoi(e));
        if (const char* e = getenv("MOT_UF_LIGHT")) h->uf_light = std::min(4096, std::max(1, atoi(e)));
        if (const char* e = getenv("MOT_UF_XBLOCKS")) h->uf_cross_blocks = std::min(32, std::max(1, atoi(e)));
        if (const char* e = getenv("MOT_UF_SPLIT")) h->uf_split = atoi(e) != 0;
        if (const char* e = getenv("MOT_UF_XMODE")) h->uf_xmode = atoi(e);
        if (const char* e = getenv("MOT_UF_PHASES")) {  // comma separated row masks; rows left out by every phase go to the last one
            unsigned seen = 0;
            int np = 0;
            for (const char* c = e; *c && np < 3;) {
                h->uf_phases[np] = (unsigned)strtoul(c, nullptr, 10) & 0x1fu & ~seen;
                seen |= h->uf_phases[np++];
                while (*c && *c != ',') ++c;
                if (*c == ',') ++c;
            }
            if (np > 0) {
                h->uf_phases[np - 1] |= 0x1fu & ~seen;
                h->uf_n_phases = np;
            }
        }
        if (const char* e = getenv("MOT_UF_FBLOCKS")) h->uf_fused_blocks = std::min(16, std::max(1, atoi(e)));
        if (const char* e = getenv("MOT_UF_WBLOCKS")) h->uf_walk_blocks = std::min(16, std::max(1, atoi(e)));
        {   // a phase lists at most (its neighbour slots) x (coarse cells <= points) tasks
            int worst = 1;
            for (int ph = 0; ph < h->uf_n_phases; ++ph) {
                int slots = 0;
                for (int row = 0; row < 5; ++row)
                    if ((h->uf_phases[ph] >> row) & 1u) slots += row == 0 ? 1 : 3;
                worst = std::max(worst, slots);
            }
            h->task_cap = (int)std::min<size_t>((size_t)worst * n + 64, 0x7ffffff0ull);
        }
        if (h->uf_xmode == 1) CK(dalloc(&h->d_tasks, (size_t)h->task_cap));
        int hb = ceil_log2(2 * (long long)n);
        if (hb < 4) hb = 4;
        h->hash_capacity = (size_t)1 << hb;
        CK(cudaMalloc(&h->d_hkeys, h->hash_capacity * 8));
        CK(dalloc(&h->d_hvals, h->hash_capacity));
        CK(dalloc(&h->d_counts, (size_t)CNT_N));
        CK(dalloc(&h->rws.hist, rs_workspace_counters()));
        CK(dalloc(&h->rws.prefix, rs_workspace_counters()));
        CK(dalloc(&h->rws.tot, (size_t)1 << RS_MAX_BITS));
        CK(dalloc(&h->rws.ghist, (size_t)OS_MAX_PASSES * ((size_t)1 << RS_MAX_BITS) + OS_MAX_PASSES));
        h->rws.status_words = (size_t)OS_MAX_PASSES * ((n + OS_TILE - 1) / OS_TILE + 1) * ((size_t)1 << RS_MAX_BITS);
        CK(dalloc(&h->rws.status, h->rws.status_words));
        h->rws.err_flag = h->d_counts + CNT_FLAGS;
        if (const char* e = getenv("MOT_SORT_MODE")) h->rws.mode = atoi(e);
        if (const char* e = getenv("MOT_SORT_BIGTILE")) h->rws.big_tile_from = atoi(e);
        if (const char* e = getenv("MOT_PLAN_SPEC")) h->plan_spec = atoi(e);
        if (const char* e = getenv("MOT_SMALL_POINTS")) h->fs_max_points = std::max(0, atoi(e));
        if (const char* e = getenv("MOT_SMALL_CELLCAP")) h->fs_cell_cap = std::max(1, atoi(e));
        CK(dalloc(&h->d_fs_state, (size_t)FS_ST_N));
        CK(dalloc(&h->d_fs_args, (size_t)1));
        CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_fs_args), sizeof(FsArgs), cudaHostAllocDefault));
        if (const char* e = getenv("MOT_SMALL_GRAPH")) h->fs_use_graph = atoi(e);
        CK(dalloc(&h->d_fs_phase_ns, (size_t)FS_PHASES));
        CK(cudaMemset(h->d_fs_phase_ns, 0, FS_PHASES * sizeof(unsigned long long)));
        {   // the small-frame kernel runs as one cluster: 16 CTAs where the device schedules them (non-portable size), else 8
            int want = 16;
            if (const char* e = getenv("MOT_SMALL_CLUSTER")) want = atoi(e);
            cudaFuncSetAttribute(k_fs_front, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaFuncSetAttribute(k_fs_tables, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaFuncSetAttribute(k_fs_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_TABLES_SMEM);
            for (int c = want; c >= 1 && h->fs_cluster == 0; c >>= 1) {
                if (c > 16) continue;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(c);
                cfg.blockDim = dim3(FS_THREADS);
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                int n_clusters = 0;
                int n2 = 0;
                if (cudaOccupancyMaxActiveClusters(&n_clusters, k_fs_front, &cfg) == cudaSuccess && n_clusters >= 1 &&
                    (cfg.dynamicSmemBytes = FS_TABLES_SMEM, cudaOccupancyMaxActiveClusters(&n2, k_fs_tables, &cfg)) == cudaSuccess && n2 >= 1)
                    h->fs_cluster = c;
            }
            cudaGetLastError();
        }
        if (const char* e = getenv("MOT_KEYS_HIST")) h->keys_hist_fused = atoi(e);
        if (const char* e = getenv("MOT_CSR_COMPACT")) h->csr_compact = atoi(e);
        if (const char* e = getenv("MOT_SORT_BITS")) h->rws.digit_bits = std::min(RS_MAX_BITS, std::max(4, atoi(e)));
        CK(dalloc(&h->d_blk, (size_t)32 * 1024));
        CK(dalloc(&h->d_bbox, (size_t)8));
        CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_pinned), 64 * sizeof(int), cudaHostAllocDefault));
        if (const char* e = getenv("MOT_HOST_STAGE")) h->stage_on = atoi(e);
        if (h->stage_on) CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_stage), STAGE_IN + STAGE_OUT, cudaHostAllocDefault));
        h->frame_capacity = 4096;
        CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_pinned_fo), (2 * h->frame_capacity + 8) * sizeof(int), cudaHostAllocDefault));
        CK(dalloc(&h->d_frame_offsets, h->frame_capacity + 2));
        CK(dalloc(&h->d_frame_offsets_in, h->frame_capacity + 2));
        CK(dalloc(&h->d_frame_stamps, h->frame_capacity + 2));
        h->c1p_tiles = (n + C1P_TILE - 1) / C1P_TILE;
        CK(dalloc(&h->d_c1p_status, h->c1p_tiles + 2));
        if (const char* e = getenv("MOT_RS_MODE")) h->rs_mode = atoi(e);
        CK(dalloc(&h->d_frame_cl_offsets, h->frame_capacity + 2));
        CK(cudaMemset(h->d_frame_offsets, 0, (h->frame_capacity + 2) * sizeof(int)));
        for (auto& e : h->ev) CK(cudaEventCreate(&e));
        for (auto& e : h->timer_ev) CK(cudaEventCreate(&e));
        h->prof.st = h->stream;
        CK(cudaFuncSetAttribute(k_uf_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(UFC_WARPS * sizeof(UfcWarpSmem))));
        CK(cudaFuncSetAttribute(k_uf_sparse2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(UFC_WARPS * sizeof(UfpWarpSmem))));
        if (const char* e = getenv("MOT_UF_MODE")) h->uf_mode = atoi(e);
        if (const char* e = getenv("MOT_UF_PAIR")) h->uf_pair = atoi(e);
        if (const char* e = getenv("MOT_UF_TMA")) h->uf_tma = atoi(e) & 1;  // 0: plain loads instead of TMA staging (A/B)
        CK(rs_configure<uint32_t>());
        CK(rs_configure<uint64_t>());
        CK(cudaFuncSetAttribute(k_clusters_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CL_SMALL_SMEM));
        CK(cudaFuncSetAttribute(k_rs_count, cudaFuncAttributeMaxDynamicSharedMemorySize, RSK_SMEM_BITMAP_MAX));
        CK(cudaFuncSetAttribute(k_rs_compact, cudaFuncAttributeMaxDynamicSharedMemorySize, RSK_SMEM_BITMAP_MAX));
        CK(cudaFuncSetAttribute(k_compact_onepass<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RSK_SMEM_BITMAP_MAX));
        if (max_tracks > 0) {
            CK(dalloc(&h->d_mstate, max_tracks * 4));
            CK(dalloc(&h->d_posvel, max_tracks * 2));
            CK(dalloc(&h->d_track_ids, max_tracks));
            CK(dalloc(&h->d_obstacles, max_tracks));
            for (int i = 0; i < 2; ++i) {
                CK(dalloc(&h->d_trk_ids[i], max_tracks));
                CK(dalloc(&h->d_trk_m[i], max_tracks * 4));
            }
            CK(dalloc(&h->d_trk_meta, (size_t)TM_N));
            CK(cudaMemset(h->d_trk_meta, 0, TM_N * sizeof(int)));
            CK(dalloc(&h->d_trk_seen, max_tracks));
            CK(dalloc(&h->d_ent_ids, max_tracks));
            CK(dalloc(&h->d_ent_slot, max_tracks));
            CK(dalloc(&h->d_ent_occ, max_tracks));
            CK(dalloc(&h->d_ent_next, max_tracks));
            CK(dalloc(&h->d_ent_prev, max_tracks));
            if (const char* e = getenv("MOT_ASSOC_FAST")) h->assoc_fast = atoi(e);
            CK(dalloc(&h->d_centroids_in, max_tracks));
        }
        return MOT_OK;
    };
    const int rc = body();
    if (rc != MOT_OK) {
        fprintf(stderr, "mot_create: %s\n", h->err.c_str());
        mot_destroy(h);
        return rc;
    }
    *out = h;
    return MOT_OK;
}

int mot_destroy(mot_handle* h) {
    if (!h) return MOT_ERR_INVALID;
    cudaSetDevice(h->device);
    if (h->stream) mot_sync(h);
    void* ptrs[] = {h->d_in, h->d_pts, h->d_spts, h->d_keys[0], h->d_keys[1], h->d_vals[0], h->d_vals[1], h->d_ckeys[0], h->d_ckeys[1],
                    h->d_croots[0], h->d_croots[1], h->d_fc_start, h->d_cc_first, h->d_parent, h->d_root, h->d_csize, h->d_cmin,
                    h->d_crank, h->d_labels, h->d_cl_offsets, h->d_hkeys, h->d_hvals, h->rws.hist, h->rws.prefix, h->rws.tot, h->rws.ghist, h->rws.status, h->d_blk,
                    h->d_counts, h->d_bbox, h->d_frame_offsets, h->d_frame_offsets_in, h->d_frame_stamps, h->d_c1p_status, h->d_frame_cl_offsets, h->d_stats, h->d_statacc, h->d_crec, h->d_dense_list, h->d_nbr, h->d_centroids, h->d_cands, h->d_bits,
                    h->d_rings, h->d_mstate, h->d_posvel, h->d_track_ids, h->d_obstacles, h->d_raw, h->d_trk_ids[0], h->d_trk_ids[1],
                    h->d_trk_rings[0], h->d_trk_rings[1], h->d_trk_m[0], h->d_trk_m[1], h->d_trk_meta, h->d_trk_seen, h->d_ent_ids, h->d_ent_slot,
                    h->d_ent_occ, h->d_ent_next, h->d_ent_prev, h->d_centroids_in, h->d_ckey, h->d_fcode, h->d_tasks, h->d_cbox, h->d_fbox, h->d_heavy1, h->d_heavy2, h->d_fs_state, h->d_fs_args, h->d_fs_phase_ns};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->h_pinned_fo) cudaFreeHost(h->h_pinned_fo);
    if (h->h_fs_args) cudaFreeHost(h->h_fs_args);
    if (h->fs_graph) cudaGraphExecDestroy(h->fs_graph);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : h->timer_ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : h->prof.pool) cudaEventDestroy(e);
    if (h->sync_ev) cudaEventDestroy(h->sync_ev);
    for (cudaEvent_t e : h->ev_uf) if (e) cudaEventDestroy(e);
    if (h->uf_stream) cudaStreamDestroy(h->uf_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MOT_OK;
}

const char* mot_last_error(mot_handle* h) { return h ? h->err.c_str() : "null handle"; }

int mot_set_map(mot_handle* h, const int8_t* occ, int width, int height, float resolution, double origin_x, double origin_y,
                const double quat_xyzw[4], int static_tolerance) {
    if (!h) return MOT_ERR_INVALID;
    if (!occ || !quat_xyzw || width <= 0 || height <= 0 || !(resolution > 0.0f) || (long long)width * height > 0x7fffffffll)
        return fail(h, MOT_ERR_INVALID, "bad map arguments");
    CK(cudaSetDevice(h->device));
    if (static_tolerance > 4) static_tolerance = 4; else if (static_tolerance < 0) static_tolerance = 0;  // MOT.cpp:95-96
    const size_t cells = (size_t)width * height;
    const size_t words = (cells + 31) / 32;
    const size_t padded_words = (words + 3) & ~(size_t)3;  // 16-byte multiple for the TMA bulk copy
    if (padded_words > h->bits_capacity_words) {
        if (h->d_bits) cudaFree(h->d_bits);
        h->d_bits = nullptr;
        CK(dalloc(&h->d_bits, padded_words));
        h->bits_capacity_words = padded_words;
    }
    int8_t* d_occ = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&d_occ), cells));
    cudaError_t e = cudaMemcpyAsync(d_occ, occ, cells, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->d_bits, 0, padded_words * 4, h->stream);
    if (e == cudaSuccess) {
        k_build_blocked_bitmap<<<(unsigned)((cells + 255) / 256), 256, 0, h->stream>>>(d_occ, width, height, static_tolerance, h->d_bits);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = mot_sync(h);
    cudaFree(d_occ);
    if (e != cudaSuccess) { h->err = std::string("mot_set_map: ") + cudaGetErrorString(e); return MOT_ERR_CUDA; }
    // yaw exactly as quaternion2eularYaw (MOT.cpp:1013-1023): double atan2 returned through a float; then the
    // float overloads of cos/sin the reference's `cos(-theta)` resolves to (MOT.cpp:677-678).
    const double siny_cosp = 2 * (quat_xyzw[3] * quat_xyzw[2] + quat_xyzw[0] * quat_xyzw[1]);
    const double cosy_cosp = 1 - 2 * (quat_xyzw[1] * quat_xyzw[1] + quat_xyzw[2] * quat_xyzw[2]);
    const float theta = (float)std::atan2(siny_cosp, cosy_cosp);
    h->mp.cs = std::cos(-theta);
    h->mp.sn = std::sin(-theta);
    h->mp.origin_x = origin_x;
    h->mp.origin_y = origin_y;
    h->mp.resolution = resolution;
    h->mp.width = width;
    h->mp.height = height;
    h->mp.n_words = (int)words;
    h->have_map = true;
    return MOT_OK;
}

int mot_set_cluster_params(mot_handle* h, float cluster_tolerance, int min_cluster_size, int max_cluster_size) {
    if (!h) return MOT_ERR_INVALID;
    if (!(cluster_tolerance > 0.0f) || !std::isfinite(cluster_tolerance)) return fail(h, MOT_ERR_INVALID, "cluster_tolerance must be > 0");
    if (min_cluster_size < 1) min_cluster_size = 1;  // PCL: a component always has >= 1 point
    if (h->tol != cluster_tolerance) h->plan_valid = false;
    h->tol = cluster_tolerance;
    h->min_size = min_cluster_size;
    h->max_size = max_cluster_size;
    return MOT_OK;
}

int mot_remove_static(mot_handle* h, const float* xyz16, size_t n, float* out_xyz16, size_t out_capacity, size_t* m) {
    int rc = check_frame_args(h, xyz16, n);
    if (rc != MOT_OK) return rc;
    if (!h->have_map) return fail(h, MOT_ERR_NO_MAP, "mot_set_map has not been called");
    if (!m) return fail(h, MOT_ERR_INVALID, "null output");
    CK(cudaSetDevice(h->device));
    *m = 0;
    if (n == 0) return MOT_OK;
    rc = reset_frame_state(h);
    if (rc != MOT_OK) return rc;
    CK(cudaMemcpyAsync(h->d_in, xyz16, n * 16, cudaMemcpyHostToDevice, h->stream));
    rc = enqueue_remove_static(h, h->d_in, (int)n, h->d_pts);
    if (rc != MOT_OK) return rc;
    CK(cudaMemcpyAsync(h->h_pinned + 8, h->d_counts, CNT_N * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(mot_sync(h));
    const size_t M = (size_t)h->h_pinned[8 + CNT_M];
    *m = M;
    if (out_xyz16) {
        if (out_capacity < M) return fail(h, MOT_ERR_CAPACITY, "output cloud buffer too small");
        if (M) CK(cudaMemcpyAsync(out_xyz16, h->d_pts, M * 16, cudaMemcpyDeviceToHost, h->stream));
        CK(mot_sync(h));
    }
    return MOT_OK;
}

// VoxelGrid downsample (reference MOT.cpp:452-456, PCL VoxelGrid restated): bbox -> fp32 voxel index -> radix sort ->
// segment heads -> point-parallel segmented mean.  d_src: cloud on the device (n points).  The result (V centroids in
// ascending voxel index) is left in h->d_pts; *m_out = V.
static int voxel_grid_device(mot_handle* h, const float4* d_src, int n, float lx, float ly, float lz, int* m_out) {
    cudaStream_t st = h->stream;
    *m_out = 0;
    if (n == 0) return MOT_OK;
    int grid = (n + 255) / 256;
    if (grid > h->num_sms * 8) grid = h->num_sms * 8;
    LAUNCH(KID_BBOX, k_bbox<<<grid, 256, 0, st>>>(d_src, n, h->d_bbox));
    CK(cudaMemcpyAsync(h->h_pinned, h->d_bbox, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(mot_sync(h));
    if (h->h_pinned[6] != 0) return fail(h, MOT_ERR_NONFINITE, "cloud contains NaN/Inf coordinates");
    VoxelParams vp{};
    const float leaf[3] = {lx, ly, lz};
    long long div[3];
    for (int d = 0; d < 3; ++d) {
        vp.inv[d] = 1.0f / leaf[d];                                                        // inverse_leaf_size_
        const float mn = ordered_to_float_bits(h->h_pinned[d]), mx = ordered_to_float_bits(h->h_pinned[3 + d]);
        vp.minb[d] = (int)std::floor(mn * vp.inv[d]);                                      // min_b_
        const int maxb = (int)std::floor(mx * vp.inv[d]);                                  // max_b_
        div[d] = (long long)maxb - vp.minb[d] + 1;                                         // div_b_
    }
    if ((double)div[0] * (double)div[1] * (double)div[2] > 2147483647.0) {  // (the int64 product itself can wrap)
        // pcl::VoxelGrid::applyFilter: "Leaf size is too small for the input dataset. Integer indices would overflow." -- PCL
        // warns and hands the input cloud back unchanged; so does this call (status MOT_WARN_VOXEL_OVERFLOW)
        CK(cudaMemcpyAsync(h->d_pts, d_src, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
        *m_out = n;
        h->err = "leaf size too small for the cloud extent: voxel index overflows, input returned unchanged (as PCL does)";
        return MOT_WARN_VOXEL_OVERFLOW;
    }
    vp.mul1 = (int)div[0];
    vp.mul2 = (int)(div[0] * div[1]);
    uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(h->d_keys[0]), reinterpret_cast<uint32_t*>(h->d_keys[1])};
    LAUNCH(KID_VOX_KEYS, k_voxel_keys<<<(n + 255) / 256, 256, 0, st>>>(d_src, n, vp, keys[0]));
    const int sb = radix_sort_pairs<uint32_t>(st, keys, h->d_vals, n, ceil_log2(div[0] * div[1] * div[2]), true, h->rws, h->prof, KID_VOX_HIST);
    const Chunking ck = make_chunking(n, 256, CELL_MAX_GRID);
    LAUNCH(KID_SEG_COUNT, k_seg_count<<<ck.grid, 256, 0, st>>>(keys[sb], n, ck.chunk, h->d_blk));
    LAUNCH(KID_SEG_WRITE, k_seg_write<<<ck.grid, 256, 0, st>>>(keys[sb], n, ck.chunk, h->d_blk, h->d_cl_offsets, h->d_counts + CNT_K));
    CK(cudaMemcpyAsync(h->h_pinned + 8, h->d_counts, CNT_N * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(mot_sync(h));
    const int V = h->h_pinned[8 + CNT_K];
    int rc = ensure_tables(h, (size_t)V, 0);
    if (rc != MOT_OK) return rc;
    LAUNCH(KID_STATS_INIT, k_stats_init<<<(V + 255) / 256, 256, 0, st>>>(h->d_statacc, V));
    const int sg = stat_groups(n, h->num_sms);
    const int per_block = 32 * sg * (STAT_THREADS / 32);
    LAUNCH(KID_STATS, k_stats_accumulate<<<(n + per_block - 1) / per_block, STAT_THREADS, 0, st>>>(d_src, h->d_cl_offsets, h->d_vals[sb], V, n, h->d_statacc, sg));
    LAUNCH(KID_VOX_FIN, k_voxel_finalize<<<(V + 255) / 256, 256, 0, st>>>(h->d_statacc, h->d_cl_offsets, V, h->d_pts));
    CK(cudaGetLastError());
    *m_out = V;
    return MOT_OK;
}

int mot_voxel_grid(mot_handle* h, const float* xyz16, size_t n, float leaf_x, float leaf_y, float leaf_z, float* out_xyz16, size_t out_capacity,
                   size_t* m) {
    int rc = check_frame_args(h, xyz16, n);
    if (rc != MOT_OK) return rc;
    if (!m || !(leaf_x > 0) || !(leaf_y > 0) || !(leaf_z > 0)) return fail(h, MOT_ERR_INVALID, "bad voxel grid arguments");
    CK(cudaSetDevice(h->device));
    *m = 0;
    if (n == 0) return MOT_OK;
    rc = reset_frame_state(h);
    if (rc != MOT_OK) return rc;
    CK(cudaMemcpyAsync(h->d_in, xyz16, n * 16, cudaMemcpyDefault, h->stream));
    int V = 0;
    rc = voxel_grid_device(h, h->d_in, (int)n, leaf_x, leaf_y, leaf_z, &V);
    if (rc < 0) return rc;
    const int warn = rc;
    *m = (size_t)V;
    if (out_xyz16) {
        if (out_capacity < (size_t)V) return fail(h, MOT_ERR_CAPACITY, "output cloud buffer too small");
        if (V) CK(cudaMemcpyAsync(out_xyz16, h->d_pts, (size_t)V * 16, cudaMemcpyDefault, h->stream));
    }
    CK(mot_sync(h));
    fold_profile(h);
    return warn;
}

// PointCloud2 wire format -> pcl::PointXYZ on the device (pcl::fromROSMsg, MOT.cpp:448-449).
int mot_unpack_pointcloud2(mot_handle* h, const uint8_t* data, size_t n_points, uint32_t point_step, uint32_t off_x, uint32_t off_y,
                           uint32_t off_z, int is_bigendian, int drop_nonfinite, float* out_xyz16, size_t out_capacity, size_t* m) {
    int rc = check_frame_args(h, data, n_points);
    if (rc != MOT_OK) return rc;
    if (!m || point_step < 12 || (uint64_t)off_x + 4 > point_step || (uint64_t)off_y + 4 > point_step || (uint64_t)off_z + 4 > point_step)
        return fail(h, MOT_ERR_INVALID, "bad PointCloud2 layout");
    if (n_points > SIZE_MAX / point_step) return fail(h, MOT_ERR_INVALID, "PointCloud2 payload size overflows");
    CK(cudaSetDevice(h->device));
    *m = 0;
    if (n_points == 0) return MOT_OK;
    rc = reset_frame_state(h);
    if (rc != MOT_OK) return rc;
    const size_t bytes = n_points * (size_t)point_step;
    rc = ensure_raw(h, bytes);
    if (rc != MOT_OK) return rc;
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->d_raw, data, bytes, cudaMemcpyDefault, st));
    const int n = (int)n_points;
    const Chunking ck = make_chunking(n, 256, RSK_MAX_GRID);
    LAUNCH(KID_PC2_UNPACK, k_pc2_unpack<<<ck.grid, 256, 0, st>>>(h->d_raw, n, ck.chunk, point_step, off_x, off_y, off_z, is_bigendian, h->d_in, h->d_blk));
    const float4* result = h->d_in;
    size_t M = n_points;
    if (drop_nonfinite) {
        const int tiles = (n + C1P_TILE - 1) / C1P_TILE;
        CK(cudaMemsetAsync(h->d_c1p_status, 0, ((size_t)tiles + 1) * sizeof(unsigned), st));
        LAUNCH(KID_PC2_COMPACT, k_compact_onepass<1><<<tiles, C1P_THREADS, 0, st>>>(h->d_in, n, h->mp, nullptr, 0, h->d_pts, h->d_c1p_status, tiles,
                                                                                 h->d_counts + CNT_M, h->d_bbox, nullptr, 1, nullptr,
                                                                                 h->d_counts + CNT_FLAGS));
        CK(cudaMemcpyAsync(h->h_pinned + 8, h->d_counts, CNT_N * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(mot_sync(h));
        M = (size_t)h->h_pinned[8 + CNT_M];
        result = h->d_pts;
    }
    CK(cudaGetLastError());
    *m = M;
    if (out_xyz16) {
        if (out_capacity < M) return fail(h, MOT_ERR_CAPACITY, "output cloud buffer too small");
        if (M) CK(cudaMemcpyAsync(out_xyz16, result, M * 16, cudaMemcpyDefault, st));
    }
    CK(mot_sync(h));
    fold_profile(h);
    return MOT_OK;
}

int mot_cluster(mot_handle* h, const float* xyz16, size_t m, int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices,
                size_t indices_capacity, int32_t* n_clusters) {
    int rc = check_frame_args(h, xyz16, m);
    if (rc != MOT_OK) return rc;
    CK(cudaSetDevice(h->device));
    begin_frame(h);
    CK(cudaEventRecord(h->ev[0], h->stream));
    rc = upload_cloud(h, h->d_pts, xyz16, m * 16);
    if (rc != MOT_OK) return rc;
    rc = frame_core(h, h->d_pts, (int)m, false, nullptr, false, 0.0);
    if (rc != MOT_OK) return rc;
    if (n_clusters) *n_clusters = h->res_K;
    rc = fetch_result(h, nullptr, 0, cluster_offsets, offsets_capacity, point_indices, indices_capacity, nullptr, nullptr, 0);
    if (rc != MOT_OK) return rc;
    return finish_timings(h);
}

int mot_cluster_stats(mot_handle* h, mot_cluster_stat* stats, size_t capacity) {
    if (!h || !stats) return h ? fail(h, MOT_ERR_INVALID, "null output") : MOT_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    return fetch_result(h, nullptr, 0, nullptr, 0, nullptr, 0, stats, nullptr, capacity);
}

int mot_get_centroid(mot_handle* h, double stamp_minus_time_init, float* out_xyzi, size_t capacity) {
    if (!h || !out_xyzi) return h ? fail(h, MOT_ERR_INVALID, "null output") : MOT_ERR_INVALID;
    if (!h->have_result) return fail(h, MOT_ERR_STATE, "no clustering result on this handle");
    if (h->res_frames != 1 && !h->res_centroids) return fail(h, MOT_ERR_STATE, "batch centroids must be requested with the batch call (indices are frame-local afterwards)");
    CK(cudaSetDevice(h->device));
    if (h->res_frames == 1) {
        int rc = compute_centroids_late(h, stamp_minus_time_init);
        if (rc != MOT_OK) return rc;
    }
    return fetch_result(h, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, out_xyzi, capacity);
}

int mot_frame_device(mot_handle* h, const float* d_xyz16, size_t n, int do_remove_static, int with_centroids, double stamp) {
    int rc = check_frame_args(h, d_xyz16, n);
    if (rc != MOT_OK) return rc;
    if (do_remove_static && !h->have_map) return fail(h, MOT_ERR_NO_MAP, "mot_set_map has not been called");
    CK(cudaSetDevice(h->device));
    begin_frame(h);
    CK(cudaEventRecord(h->ev[0], h->stream));
    const float4* src = reinterpret_cast<const float4*>(d_xyz16);
    rc = frame_core(h, src, (int)n, do_remove_static != 0, h->d_pts, with_centroids != 0, stamp);
    if (rc != MOT_OK) return rc;
    return finish_timings(h);
}

int mot_frame(mot_handle* h, const float* xyz16, size_t n, double stamp, float* kept_xyz16, size_t kept_capacity, size_t* m,
              int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters,
              mot_cluster_stat* stats, float* centroids_xyzi, size_t table_capacity) {
    int rc = check_frame_args(h, xyz16, n);
    if (rc != MOT_OK) return rc;
    if (!h->have_map) return fail(h, MOT_ERR_NO_MAP, "mot_set_map has not been called");
    CK(cudaSetDevice(h->device));
    begin_frame(h);
    CK(cudaEventRecord(h->ev[0], h->stream));
    rc = upload_cloud(h, h->d_in, xyz16, n * 16);
    if (rc != MOT_OK) return rc;
    rc = frame_core(h, h->d_in, (int)n, true, h->d_pts, centroids_xyzi != nullptr, stamp);
    if (rc != MOT_OK) return rc;
    if (m) *m = (size_t)h->res_M;
    if (n_clusters) *n_clusters = h->res_K;
    rc = fetch_result(h, kept_xyz16, kept_capacity, cluster_offsets, offsets_capacity, point_indices, indices_capacity, stats, centroids_xyzi,
                      table_capacity);
    if (rc != MOT_OK) return rc;
    return finish_timings(h);
}

int mot_result_counts(mot_handle* h, size_t* m, int32_t* n_clusters, size_t* n_indices) {
    if (!h) return MOT_ERR_INVALID;
    if (!h->have_result) return fail(h, MOT_ERR_STATE, "no clustering result on this handle");
    if (m) *m = (size_t)h->res_M;
    if (n_clusters) *n_clusters = h->res_K;
    if (n_indices) *n_indices = (size_t)h->res_total;
    return MOT_OK;
}

int mot_result_grid(mot_handle* h, int32_t* fine_cells, int32_t* coarse_cells, int32_t* key_bits) {
    if (!h) return MOT_ERR_INVALID;
    if (!h->have_result) return fail(h, MOT_ERR_STATE, "no clustering result on this handle");
    if (fine_cells) *fine_cells = h->res_fine;
    if (coarse_cells) *coarse_cells = h->res_coarse;
    if (key_bits) *key_bits = h->res_key_bits;
    return MOT_OK;
}

int mot_result_counters(mot_handle* h, int32_t* out, int capacity) {
    if (!h || !out || capacity < 1) return MOT_ERR_INVALID;
    if (!h->have_result) return fail(h, MOT_ERR_STATE, "no clustering result on this handle");
    for (int i = 0; i < capacity; ++i) out[i] = i < CNT_N ? h->res_counters[i] : 0;
    return MOT_OK;
}

int mot_grid_plan(mot_handle* h, int enable, int* hits, int* misses) {
    if (!h) return MOT_ERR_INVALID;
    if (enable >= 0) {
        h->plan_spec = enable ? 1 : 0;
        h->plan_valid = false;
    }
    if (hits) *hits = h->plan_hits;
    if (misses) *misses = h->plan_misses;
    return MOT_OK;
}

int mot_small_frames(mot_handle* h, int max_points, int* hits, int* misses) {
    if (!h) return MOT_ERR_INVALID;
    if (max_points >= 0) {
        h->fs_max_points = max_points;
        h->fs_skip = 0;
    }
    if (hits) *hits = h->fs_hits;
    if (misses) *misses = h->fs_misses;
    return h->fs_cluster;
}

int mot_small_frame_phases(mot_handle* h, uint64_t* ns, int capacity) {
    if (!h || !ns || capacity < 1) return MOT_ERR_INVALID;
    unsigned long long t[FS_PHASES];
    CK(cudaSetDevice(h->device));
    CK(mot_sync(h));
    CK(cudaMemcpy(t, h->d_fs_phase_ns, sizeof(t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < capacity; ++i) ns[i] = i < FS_PHASES ? (uint64_t)t[i] : 0;
    return MOT_OK;
}

int mot_debug_stats(mot_handle* h, uint64_t* out, int capacity) {
    if (!h || !out || capacity < 1) return MOT_ERR_INVALID;
    for (int i = 0; i < capacity; ++i) out[i] = 0;
#ifdef MOT_UF_STATS
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    unsigned long long tmp[ST_N];
    CK(cudaMemcpyFromSymbol(tmp, g_uf_stats, sizeof(tmp)));
    for (int i = 0; i < capacity && i < ST_N; ++i) out[i] = tmp[i];
    std::memset(tmp, 0, sizeof(tmp));
    CK(cudaMemcpyToSymbol(g_uf_stats, tmp, sizeof(tmp)));
    return 1;
#else
    return MOT_OK;
#endif
}

int mot_result_device_ptrs(mot_handle* h, const float** d_kept, const int32_t** d_offsets, const int32_t** d_indices,
                           const mot_cluster_stat** d_stats, const float** d_centroids) {
    if (!h) return MOT_ERR_INVALID;
    if (!h->have_result) return fail(h, MOT_ERR_STATE, "no clustering result on this handle");
    if (d_kept) *d_kept = reinterpret_cast<const float*>(h->res_cloud);
    if (d_offsets) *d_offsets = h->d_cl_offsets;
    if (d_indices) *d_indices = reinterpret_cast<const int32_t*>(h->d_vals[h->res_idx_buf]);
    if (d_stats) *d_stats = reinterpret_cast<const mot_cluster_stat*>(h->d_stats);
    if (d_centroids) *d_centroids = h->res_centroids ? reinterpret_cast<const float*>(h->d_centroids) : nullptr;
    return MOT_OK;
}

int mot_result_fetch(mot_handle* h, float* kept_xyz16, size_t kept_capacity, int32_t* cluster_offsets, size_t offsets_capacity,
                     int32_t* point_indices, size_t indices_capacity, mot_cluster_stat* stats, float* centroids_xyzi, size_t table_capacity) {
    if (!h) return MOT_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    return fetch_result(h, kept_xyz16, kept_capacity, cluster_offsets, offsets_capacity, point_indices, indices_capacity, stats, centroids_xyzi,
                        table_capacity);
}

int mot_result_labels(mot_handle* h, int32_t* labels, size_t capacity) {
    if (!h || !labels) return h ? fail(h, MOT_ERR_INVALID, "null output") : MOT_ERR_INVALID;
    if (!h->have_result) return fail(h, MOT_ERR_STATE, "no clustering result on this handle");
    if (capacity < (size_t)h->res_M) return fail(h, MOT_ERR_CAPACITY, "labels buffer too small");
    CK(cudaSetDevice(h->device));
    if (h->res_M) CK(cudaMemcpyAsync(labels, h->d_labels, (size_t)h->res_M * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(mot_sync(h));
    return MOT_OK;
}

int mot_last_timings(mot_handle* h, mot_timings* t) {
    if (!h || !t) return MOT_ERR_INVALID;
    *t = h->tim;
    return MOT_OK;
}

int mot_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return MOT_ERR_INVALID;
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? MOT_OK : MOT_ERR_CUDA;
}
int mot_host_unregister(void* ptr) {
    if (!ptr) return MOT_ERR_INVALID;
    return cudaHostUnregister(ptr) == cudaSuccess ? MOT_OK : MOT_ERR_CUDA;
}

int mot_last_launches(mot_handle* h) { return h ? h->prof.launches : MOT_ERR_INVALID; }

int mot_set_profiling(mot_handle* h, int on) {
    if (!h) return MOT_ERR_INVALID;
    h->prof.on = on != 0;
    std::memset(h->prof_ms, 0, sizeof(h->prof_ms));
    std::memset(h->prof_n, 0, sizeof(h->prof_n));
    return MOT_OK;
}
int mot_profile_kernels(void) { return KID_N; }
const char* mot_profile_kernel_name(int kid) { return kid >= 0 && kid < KID_N ? kKernelNames[kid] : ""; }
int mot_profile_read(mot_handle* h, float* ms_total, int32_t* launches, int capacity) {
    if (!h || !ms_total || !launches || capacity < KID_N) return MOT_ERR_INVALID;
    for (int i = 0; i < KID_N; ++i) { ms_total[i] = h->prof_ms[i]; launches[i] = h->prof_n[i]; }
    return MOT_OK;
}

int mot_timer_start(mot_handle* h) {
    if (!h) return MOT_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->timer_ev[0], h->stream));
    return MOT_OK;
}
int mot_timer_stop(mot_handle* h, float* ms) {
    if (!h || !ms) return MOT_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());  // the stopwatch covers every handle / stream of this process on the device
    CK(cudaEventRecord(h->timer_ev[1], h->stream));
    CK(cudaEventSynchronize(h->timer_ev[1]));
    CK(cudaEventElapsedTime(ms, h->timer_ev[0], h->timer_ev[1]));
    return MOT_OK;
}

// ---- batches -------------------------------------------------------------------------------------------------
// frame_offsets -> device (d_frame_offsets_in when removeStatic will remap them, else d_frame_offsets), optional stamps
static int batch_setup(mot_handle* h, const int64_t* frame_offsets, int n_frames, size_t* total, bool remap, const float* frame_stamps) {
    if (!h) return MOT_ERR_INVALID;
    if (!frame_offsets || n_frames < 1) return fail(h, MOT_ERR_INVALID, "bad frame_offsets");
    if ((size_t)n_frames > h->frame_capacity) return fail(h, MOT_ERR_CAPACITY, "too many frames in one batch (max 4096)");
    if (frame_offsets[0] != 0) return fail(h, MOT_ERR_INVALID, "frame_offsets[0] must be 0");
    if (remap && !h->have_map) return fail(h, MOT_ERR_NO_MAP, "mot_set_map has not been called");
    int* fo = h->h_pinned_fo;
    for (int f = 0; f <= n_frames; ++f) {
        if (f && frame_offsets[f] < frame_offsets[f - 1]) return fail(h, MOT_ERR_INVALID, "frame_offsets must be non-decreasing");
        if (frame_offsets[f] > (int64_t)h->max_points) return fail(h, MOT_ERR_CAPACITY, "batch larger than the handle's max_points");
        fo[f] = (int)frame_offsets[f];
    }
    *total = (size_t)frame_offsets[n_frames];
    CK(cudaSetDevice(h->device));
    int rc = reset_frame_state(h);
    if (rc != MOT_OK) return rc;
    CK(cudaMemcpyAsync(remap ? h->d_frame_offsets_in : h->d_frame_offsets, fo, (size_t)(n_frames + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->have_frame_stamps = frame_stamps != nullptr;
    if (frame_stamps) {
        float* fs = reinterpret_cast<float*>(h->h_pinned_fo + h->frame_capacity + 2);
        std::memcpy(fs, frame_stamps, (size_t)n_frames * sizeof(float));
        CK(cudaMemcpyAsync(h->d_frame_stamps, fs, (size_t)n_frames * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    }
    return MOT_OK;  // the pinned staging is reused only after the next synchronising call on this handle
}

// core of every batch entry: cloud on the device (d_src, total points); removeStatic optional; results stay on the device
static int batch_core(mot_handle* h, const float4* d_src, size_t total, int n_frames, int do_remove_static, int with_centroids) {
    int rc;
    if (do_remove_static && total > 0) {
        if (d_src == h->d_pts) return fail(h, MOT_ERR_STATE, "internal: removeStatic source aliases its destination");
        rc = enqueue_remove_static(h, d_src, (int)total, h->d_pts, n_frames);
        if (rc != MOT_OK) return rc;
        rc = cluster_core(h, h->d_pts, -1, n_frames, with_centroids != 0, 0.0);
    } else {
        if (do_remove_static && n_frames > 1)  // nothing to compact: the kept cloud's frame boundaries are all zero
            CK(cudaMemsetAsync(h->d_frame_offsets, 0, (size_t)(n_frames + 1) * sizeof(int), h->stream));
        rc = cluster_core(h, d_src, (int)total, n_frames, with_centroids != 0, 0.0);
    }
    return rc;
}

int mot_frame_batch_device(mot_handle* h, const float* d_xyz16, const int64_t* frame_offsets, int n_frames, int do_remove_static,
                           int with_centroids, const float* frame_stamps) {
    size_t total = 0;
    int rc = batch_setup(h, frame_offsets, n_frames, &total, do_remove_static != 0, frame_stamps);
    if (rc != MOT_OK) return rc;
    if (total > 0 && !d_xyz16) return fail(h, MOT_ERR_INVALID, "null point buffer");
    CK(cudaEventRecord(h->ev[0], h->stream));
    rc = batch_core(h, reinterpret_cast<const float4*>(d_xyz16), total, n_frames, do_remove_static, with_centroids);
    if (rc != MOT_OK) return rc;
    return finish_timings(h);
}

int mot_cluster_batch_device(mot_handle* h, const float* d_xyz16, const int64_t* frame_offsets, int n_frames) {
    return mot_frame_batch_device(h, d_xyz16, frame_offsets, n_frames, 0, 0, nullptr);
}

// copies the batch result to caller buffers (host or device pointers)
static int batch_fetch(mot_handle* h, int n_frames, int32_t* frame_kept_offsets, int32_t* frame_cluster_offsets, int32_t* cluster_offsets,
                       size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters, mot_cluster_stat* stats,
                       float* centroids_xyzi, size_t table_capacity) {
    if (n_clusters) *n_clusters = h->res_K;
    if (n_frames > 1) {
        if (frame_kept_offsets) CK(cudaMemcpyAsync(frame_kept_offsets, h->d_frame_offsets, (size_t)(n_frames + 1) * 4, cudaMemcpyDefault, h->stream));
        if (frame_cluster_offsets) CK(cudaMemcpyAsync(frame_cluster_offsets, h->d_frame_cl_offsets, (size_t)(n_frames + 1) * 4, cudaMemcpyDefault, h->stream));
    } else {
        const int32_t kept[2] = {0, h->res_M}, cl[2] = {0, h->res_K};
        if (frame_kept_offsets) CK(cudaMemcpyAsync(frame_kept_offsets, kept, sizeof(kept), cudaMemcpyDefault, h->stream));
        if (frame_cluster_offsets) CK(cudaMemcpyAsync(frame_cluster_offsets, cl, sizeof(cl), cudaMemcpyDefault, h->stream));
        CK(mot_sync(h));  // the two arrays live on this stack frame
    }
    return fetch_result(h, nullptr, 0, cluster_offsets, offsets_capacity, point_indices, indices_capacity, stats, centroids_xyzi, table_capacity);
}

int mot_frame_batch(mot_handle* h, const float* xyz, int point_stride_bytes, const int64_t* frame_offsets, int n_frames, int do_remove_static,
                    const float* frame_stamps, int32_t* frame_kept_offsets, int32_t* frame_cluster_offsets, int32_t* cluster_offsets,
                    size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters, mot_cluster_stat* stats,
                    float* centroids_xyzi, size_t table_capacity) {
    if (h && point_stride_bytes != 16 && point_stride_bytes != 12) return fail(h, MOT_ERR_INVALID, "point_stride_bytes must be 16 (pcl::PointXYZ) or 12 (packed xyz)");
    size_t total = 0;
    int rc = batch_setup(h, frame_offsets, n_frames, &total, do_remove_static != 0, frame_stamps);
    if (rc != MOT_OK) return rc;
    if (total > 0 && !xyz) return fail(h, MOT_ERR_INVALID, "null point buffer");
    cudaStream_t st = h->stream;
    CK(cudaEventRecord(h->ev[0], st));
    // the cloud lands in d_in when removeStatic will compact it into d_pts, else directly in d_pts
    float4* d_cloud = do_remove_static ? h->d_in : h->d_pts;
    if (total) {
        if (point_stride_bytes == 16) {
            CK(cudaMemcpyAsync(d_cloud, xyz, total * 16, cudaMemcpyHostToDevice, st));
        } else {  // packed xyz: 25 % fewer PCIe bytes, expanded to the PointXYZ layout on the device (k_pc2_unpack)
            rc = ensure_raw(h, total * 12);
            if (rc != MOT_OK) return rc;
            CK(cudaMemcpyAsync(h->d_raw, xyz, total * 12, cudaMemcpyHostToDevice, st));
            const Chunking ck = make_chunking((long long)total, 256, RSK_MAX_GRID);
            LAUNCH(KID_PC2_UNPACK, k_pc2_unpack<<<ck.grid, 256, 0, st>>>(h->d_raw, (int)total, ck.chunk, 12u, 0u, 4u, 8u, 0, d_cloud, h->d_blk));
        }
    }
    rc = batch_core(h, d_cloud, total, n_frames, do_remove_static, centroids_xyzi != nullptr);
    if (rc != MOT_OK) return rc;
    rc = batch_fetch(h, n_frames, frame_kept_offsets, frame_cluster_offsets, cluster_offsets, offsets_capacity, point_indices, indices_capacity,
                     n_clusters, stats, centroids_xyzi, table_capacity);
    if (rc != MOT_OK) return rc;
    return finish_timings(h);
}

int mot_cluster_batch(mot_handle* h, const float* xyz16, const int64_t* frame_offsets, int n_frames, int32_t* frame_cluster_offsets,
                      int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters) {
    return mot_frame_batch(h, xyz16, 16, frame_offsets, n_frames, 0, nullptr, nullptr, frame_cluster_offsets, cluster_offsets, offsets_capacity,
                           point_indices, indices_capacity, n_clusters, nullptr, nullptr, 0);
}

// ---- fused clusterPointCloud (MOT.cpp:444-505): fromROSMsg -> VoxelGrid -> removeStatic -> extract -> getCentroid ----------------
int mot_cluster_pointcloud2(mot_handle* h, const uint8_t* data, size_t n_points, uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                            int is_bigendian, float voxel_leaf_size, int do_remove_static, double stamp_minus_time_init, float* kept_xyz16,
                            size_t kept_capacity, size_t* m, int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices,
                            size_t indices_capacity, int32_t* n_clusters, mot_cluster_stat* stats, float* centroids_xyzi, size_t table_capacity) {
    int rc = check_frame_args(h, data, n_points);
    if (rc != MOT_OK) return rc;
    if (point_step < 12 || (uint64_t)off_x + 4 > point_step || (uint64_t)off_y + 4 > point_step || (uint64_t)off_z + 4 > point_step)
        return fail(h, MOT_ERR_INVALID, "bad PointCloud2 layout");
    if (n_points > SIZE_MAX / point_step) return fail(h, MOT_ERR_INVALID, "PointCloud2 payload size overflows");
    if (do_remove_static && !h->have_map) return fail(h, MOT_ERR_NO_MAP, "mot_set_map has not been called");
    CK(cudaSetDevice(h->device));
    rc = reset_frame_state(h);
    if (rc != MOT_OK) return rc;
    cudaStream_t st = h->stream;
    CK(cudaEventRecord(h->ev[0], st));
    int warn = MOT_OK;
    int n = (int)n_points;
    const float4* cloud = h->d_pts;
    if (n > 0) {
        // fromROSMsg: wire format -> PointXYZ (d_in), then the finite points in input order (d_pts) -- PCL's VoxelGrid / kd-tree
        // skip non-finite points of a non-dense cloud
        rc = ensure_raw(h, n_points * (size_t)point_step);
        if (rc != MOT_OK) return rc;
        CK(cudaMemcpyAsync(h->d_raw, data, n_points * (size_t)point_step, cudaMemcpyDefault, st));
        const Chunking ck = make_chunking(n, 256, RSK_MAX_GRID);
        LAUNCH(KID_PC2_UNPACK, k_pc2_unpack<<<ck.grid, 256, 0, st>>>(h->d_raw, n, ck.chunk, point_step, off_x, off_y, off_z, is_bigendian, h->d_in, h->d_blk));
        const int tiles = (n + C1P_TILE - 1) / C1P_TILE;
        CK(cudaMemsetAsync(h->d_c1p_status, 0, ((size_t)tiles + 1) * sizeof(unsigned), st));
        LAUNCH(KID_PC2_COMPACT, k_compact_onepass<1><<<tiles, C1P_THREADS, 0, st>>>(h->d_in, n, h->mp, nullptr, 0, h->d_pts, h->d_c1p_status, tiles,
                                                                                 h->d_counts + CNT_M, h->d_bbox, nullptr, 1, nullptr,
                                                                                 h->d_counts + CNT_FLAGS));
        CK(cudaMemcpyAsync(h->h_pinned + 8, h->d_counts, CNT_N * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(mot_sync(h));
        n = h->h_pinned[8 + CNT_M];
        if (voxel_leaf_size > 0.0f && n > 0) {  // vg.setLeafSize(L, L, 20 L); vg.filter() (MOT.cpp:452-456); d_pts -> d_pts
            rc = reset_bbox(h);
            if (rc != MOT_OK) return rc;
            int V = 0;
            rc = voxel_grid_device(h, h->d_pts, n, voxel_leaf_size, voxel_leaf_size, 20.0f * voxel_leaf_size, &V);
            if (rc < 0) return rc;
            warn = rc;
            n = V;
        }
    }
    // removeStatic (MOT.cpp:461): d_pts -> d_in, then extract + getCentroid (frame_core clears the counters and the bounding box again)
    rc = frame_core(h, cloud, n, do_remove_static && n > 0, h->d_in, centroids_xyzi != nullptr, stamp_minus_time_init);
    if (rc != MOT_OK) return rc;
    if (m) *m = (size_t)h->res_M;
    if (n_clusters) *n_clusters = h->res_K;
    rc = fetch_result(h, kept_xyz16, kept_capacity, cluster_offsets, offsets_capacity, point_indices, indices_capacity, stats, centroids_xyzi,
                      table_capacity);
    if (rc != MOT_OK) return rc;
    rc = finish_timings(h);
    return rc != MOT_OK ? rc : warn;
}

// ---- multi-GPU frame batches: one host thread per handle, contiguous frame ranges, tables land in the caller's buffers ----
namespace {
__global__ void k_add_offset(int* __restrict__ a, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += v;
}
struct ShardJob {
    mot_handle* h = nullptr;
    int f0 = 0, f1 = 0;  // frame range
    int rc = MOT_OK;
    int K = 0, total = 0, M = 0;
    int k_base = 0, idx_base = 0, kept_base = 0;
};
}  // namespace

int mot_batch_run(mot_handle* const* handles, int n_handles, const float* xyz, int point_stride_bytes, const int64_t* frame_offsets, int n_frames,
                  int do_remove_static, const float* frame_stamps, int32_t* frame_kept_offsets, int32_t* frame_cluster_offsets,
                  int32_t* cluster_offsets, size_t offsets_capacity, int32_t* point_indices, size_t indices_capacity, int32_t* n_clusters,
                  mot_cluster_stat* stats, float* centroids_xyzi, size_t table_capacity) {
    if (!handles || n_handles < 1 || !handles[0]) return MOT_ERR_INVALID;
    mot_handle* h0 = handles[0];
    if (!frame_offsets || n_frames < 1 || frame_offsets[0] != 0) return fail(h0, MOT_ERR_INVALID, "bad frame_offsets");
    if (point_stride_bytes != 16 && point_stride_bytes != 12) return fail(h0, MOT_ERR_INVALID, "point_stride_bytes must be 16 or 12");
    for (int g = 0; g < n_handles; ++g)
        if (!handles[g]) return fail(h0, MOT_ERR_INVALID, "null handle");
    std::vector<ShardJob> jobs((size_t)n_handles);
    for (int g = 0; g < n_handles; ++g) {  // contiguous frame ranges, the rule of shard.frame_range
        jobs[g].h = handles[g];
        jobs[g].f0 = (int)((long long)n_frames * g / n_handles);
        jobs[g].f1 = (int)((long long)n_frames * (g + 1) / n_handles);
    }
    const size_t stride_floats = (size_t)point_stride_bytes / 4;
    // phase 1: every handle runs its frame range, results stay on its device
    auto phase1 = [&](ShardJob& j) {
        const int nf = j.f1 - j.f0;
        if (nf == 0) return;
        std::vector<int64_t> fo((size_t)nf + 1);
        for (int f = 0; f <= nf; ++f) fo[f] = frame_offsets[j.f0 + f] - frame_offsets[j.f0];
        mot_handle* h = j.h;
        size_t total = 0;
        j.rc = batch_setup(h, fo.data(), nf, &total, do_remove_static != 0, frame_stamps ? frame_stamps + j.f0 : nullptr);
        if (j.rc != MOT_OK) return;
        j.rc = [&]() -> int {
            cudaStream_t st = h->stream;
            CK(cudaEventRecord(h->ev[0], st));
            float4* d_cloud = do_remove_static ? h->d_in : h->d_pts;
            const float* src = xyz + (size_t)frame_offsets[j.f0] * stride_floats;
            if (total) {
                if (point_stride_bytes == 16) {
                    CK(cudaMemcpyAsync(d_cloud, src, total * 16, cudaMemcpyDefault, st));
                } else {
                    int rc = ensure_raw(h, total * 12);
                    if (rc != MOT_OK) return rc;
                    CK(cudaMemcpyAsync(h->d_raw, src, total * 12, cudaMemcpyDefault, st));
                    const Chunking ck = make_chunking((long long)total, 256, RSK_MAX_GRID);
                    LAUNCH(KID_PC2_UNPACK, k_pc2_unpack<<<ck.grid, 256, 0, st>>>(h->d_raw, (int)total, ck.chunk, 12u, 0u, 4u, 8u, 0, d_cloud, h->d_blk));
                }
            }
            int rc = batch_core(h, d_cloud, total, nf, do_remove_static, centroids_xyzi != nullptr);
            if (rc != MOT_OK) return rc;
            rc = finish_timings(h);
            return rc;
        }();
        j.K = h->res_K; j.total = h->res_total; j.M = h->res_M;
    };
    // phase 2: shift the shard's offsets by what precedes it and copy its tables into the caller's arrays
    auto phase2 = [&](ShardJob& j) {
        const int nf = j.f1 - j.f0;
        if (nf == 0) return;
        mot_handle* h = j.h;
        j.rc = [&]() -> int {
            CK(cudaSetDevice(h->device));
            cudaStream_t st = h->stream;
            const bool last = j.f1 == n_frames;
            if (j.idx_base) k_add_offset<<<(j.K + 1 + 255) / 256, 256, 0, st>>>(h->d_cl_offsets, j.K + 1, j.idx_base);
            if (cluster_offsets) CK(cudaMemcpyAsync(cluster_offsets + j.k_base, h->d_cl_offsets, (size_t)(j.K + (last ? 1 : 0)) * 4, cudaMemcpyDefault, st));
            if (point_indices && j.total) CK(cudaMemcpyAsync(point_indices + j.idx_base, h->d_vals[h->res_idx_buf], (size_t)j.total * 4, cudaMemcpyDefault, st));
            if (stats && j.K) CK(cudaMemcpyAsync(stats + j.k_base, h->d_stats, (size_t)j.K * sizeof(ClusterStat), cudaMemcpyDefault, st));
            if (centroids_xyzi && j.K) CK(cudaMemcpyAsync(centroids_xyzi + 4 * (size_t)j.k_base, h->d_centroids, (size_t)j.K * 16, cudaMemcpyDefault, st));
            if (nf > 1) {
                if (frame_cluster_offsets) {
                    if (j.k_base) k_add_offset<<<(nf + 1 + 255) / 256, 256, 0, st>>>(h->d_frame_cl_offsets, nf + 1, j.k_base);
                    CK(cudaMemcpyAsync(frame_cluster_offsets + j.f0, h->d_frame_cl_offsets, (size_t)(nf + (last ? 1 : 0)) * 4, cudaMemcpyDefault, st));
                }
                if (frame_kept_offsets) {
                    if (j.kept_base) k_add_offset<<<(nf + 1 + 255) / 256, 256, 0, st>>>(h->d_frame_offsets, nf + 1, j.kept_base);
                    CK(cudaMemcpyAsync(frame_kept_offsets + j.f0, h->d_frame_offsets, (size_t)(nf + (last ? 1 : 0)) * 4, cudaMemcpyDefault, st));
                }
            } else {  // a single frame has no per-frame tables on the device
                const int32_t cl[2] = {j.k_base, j.k_base + j.K}, kept[2] = {j.kept_base, j.kept_base + j.M};
                if (frame_cluster_offsets) CK(cudaMemcpyAsync(frame_cluster_offsets + j.f0, cl, (last ? 2 : 1) * 4, cudaMemcpyDefault, st));
                if (frame_kept_offsets) CK(cudaMemcpyAsync(frame_kept_offsets + j.f0, kept, (last ? 2 : 1) * 4, cudaMemcpyDefault, st));
                CK(mot_sync(h));
            }
            CK(cudaGetLastError());
            CK(mot_sync(h));
            h->have_result = false;  // the shard's tables were shifted in place: they now belong to the merged result
            return MOT_OK;
        }();
    };
    auto run_all = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int g = 1; g < n_handles; ++g) th.emplace_back([&, g] { fn(jobs[g]); });
        fn(jobs[0]);
        for (auto& t : th) t.join();
    };
    run_all(phase1);
    for (auto& j : jobs)
        if (j.rc != MOT_OK) { if (j.h != h0) h0->err = j.h->err; return j.rc; }
    long long k_sum = 0, idx_sum = 0, kept_sum = 0;
    for (auto& j : jobs) {
        j.k_base = (int)k_sum; j.idx_base = (int)idx_sum; j.kept_base = (int)kept_sum;
        k_sum += j.K; idx_sum += j.total; kept_sum += j.M;
    }
    if (n_clusters) *n_clusters = (int32_t)k_sum;
    if (idx_sum > 0x7fffffffll || kept_sum > 0x7fffffffll) return fail(h0, MOT_ERR_CAPACITY, "batch result does not fit 32-bit offsets");
    if (cluster_offsets && offsets_capacity < (size_t)k_sum + 1) return fail(h0, MOT_ERR_CAPACITY, "cluster_offsets buffer too small");
    if (point_indices && indices_capacity < (size_t)idx_sum) return fail(h0, MOT_ERR_CAPACITY, "point_indices buffer too small");
    if ((stats || centroids_xyzi) && table_capacity < (size_t)k_sum) return fail(h0, MOT_ERR_CAPACITY, "table buffers too small");
    run_all(phase2);
    for (auto& j : jobs)
        if (j.rc != MOT_OK) { if (j.h != h0) h0->err = j.h->err; return j.rc; }
    return MOT_OK;
}

// ---- IHGP ----------------------------------------------------------------------------------------------------
namespace {
struct M2 { double a, b, c, d; };
M2 mul2(const M2& x, const M2& y) { return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; }
M2 tr2(const M2& x) { return {x.a, x.c, x.b, x.d}; }
M2 add2(const M2& x, const M2& y) { return {x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
M2 sub2(const M2& x, const M2& y) { return {x.a - y.a, x.b - y.b, x.c - y.c, x.d - y.d}; }

// Host-side constant setup: Matern-3/2 model (M32.cpp:15-24), discretisation, fixed-point DARE
// (IHGP.cpp:213-252, <= 100 iterations, Frobenius 1e-10), stationary gain (IHGP.cpp:27-37) and the RTS smoother
// gain (IHGP.cpp:168-170).  2x2 closed forms; no Eigen.
void ihgp_setup_axis(double dt, const double hyp[3], double out[16]) {
    const double sigma2 = hyp[0], magn = hyp[1], ell = hyp[2];
    const double lam = std::sqrt(3.0) / ell;
    const M2 Pinf{magn, 0, 0, magn * lam * lam};
    const double R = sigma2;
    const double ex = std::exp(-lam * dt);  // expm(F dt) for the double eigenvalue -lam
    const M2 A{ex * (1 + lam * dt), ex * dt, ex * (-lam * lam * dt), ex * (1 - lam * dt)};
    const M2 Q = sub2(Pinf, mul2(mul2(A, Pinf), tr2(A)));
    M2 X{1, 0, 0, 1};
    for (int n = 0; n < 100; ++n) {
        const M2 Xp = X;
        double k0 = 0, k1 = 0;
        if (!(std::fabs(R) < 1e-15)) {
            const double s = X.a + R;
            const double v0 = X.a / s, v1 = X.c / s;
            k0 = A.a * v0 + A.b * v1;
            k1 = A.c * v0 + A.d * v1;
        }
        const M2 AKB{A.a - k0, A.b, A.c - k1, A.d};
        const M2 KRK{k0 * R * k0, k0 * R * k1, k1 * R * k0, k1 * R * k1};
        X = add2(add2(mul2(mul2(AKB, X), tr2(AKB)), KRK), Q);
        const M2 dX = sub2(X, Xp);
        if (std::sqrt(dX.a * dX.a + dX.b * dX.b + dX.c * dX.c + dX.d * dX.d) < 1e-10) break;
    }
    const M2 PP = X;
    const double S = PP.a + R;
    const double K0 = PP.a / S, K1 = PP.c / S;
    const M2 PF = sub2(PP, M2{K0 * PP.a, K0 * PP.b, K1 * PP.a, K1 * PP.b});
    const M2 AKHA = sub2(A, M2{K0 * A.a, K0 * A.b, K1 * A.a, K1 * A.b});
    const M2 APF = mul2(A, PF);
    const M2 PPs = add2(mul2(APF, tr2(A)), Q);
    const double det = PPs.a * PPs.d - PPs.b * PPs.c;
    const M2 inv{PPs.d / det, -PPs.b / det, -PPs.c / det, PPs.a / det};
    const M2 G = tr2(mul2(inv, APF));
    const double o[16] = {A.a, A.b, A.c, A.d, AKHA.a, AKHA.b, AKHA.c, AKHA.d, K0, K1, G.a, G.b, G.c, G.d, S, lam};
    std::memcpy(out, o, sizeof(o));
}
}  // namespace

int mot_ihgp_configure(mot_handle* h, double dt, float lpf_tau, const double hyp_x[3], const double hyp_y[3], int data_length) {
    if (!h) return MOT_ERR_INVALID;
    if (!hyp_x || !hyp_y || !(dt > 0) || data_length < 3 || data_length > 1024) return fail(h, MOT_ERR_INVALID, "bad IHGP configuration");
    for (int a = 0; a < 2; ++a) {
        const double* hyp = a == 0 ? hyp_x : hyp_y;
        if (!(hyp[0] > 0) || !(hyp[1] > 0) || !(hyp[2] > 0)) return fail(h, MOT_ERR_INVALID, "IHGP hyper-parameters must be positive");
        ihgp_setup_axis((double)(float)dt, hyp, h->ihgp_consts[a]);  // dt_gp is a float in the reference (MOT.h:113) and widens at the ctor call
        std::memcpy(h->ihgp_axis[a].A, h->ihgp_consts[a] + 0, 4 * sizeof(double));
        std::memcpy(h->ihgp_axis[a].AKHA, h->ihgp_consts[a] + 4, 4 * sizeof(double));
        std::memcpy(h->ihgp_axis[a].K, h->ihgp_consts[a] + 8, 2 * sizeof(double));
        std::memcpy(h->ihgp_axis[a].G, h->ihgp_consts[a] + 10, 4 * sizeof(double));
    }
    h->ihgp_dt = dt;
    h->ihgp_tau = lpf_tau;
    h->ihgp_L = data_length;
    h->ihgp_ready = true;
    return MOT_OK;
}

int mot_ihgp_constants(mot_handle* h, int axis, double* consts16) {
    if (!h || !consts16 || axis < 0 || axis > 1) return MOT_ERR_INVALID;
    if (!h->ihgp_ready) return fail(h, MOT_ERR_STATE, "mot_ihgp_configure has not been called");
    std::memcpy(consts16, h->ihgp_consts[axis], 16 * sizeof(double));
    return MOT_OK;
}

static int ihgp_step_impl(mot_handle* h, const float* rings, int n_tracks, const int32_t* ids, double* m_state, float* pos_vel,
                          mot_obstacle* obstacles);

int mot_ihgp_step(mot_handle* h, const float* rings, int n_tracks, double* m_state, float* pos_vel) {
    return ihgp_step_impl(h, rings, n_tracks, nullptr, m_state, pos_vel, nullptr);
}

int mot_ihgp_step_obstacles(mot_handle* h, const float* rings, int n_tracks, const int32_t* track_ids, double* m_state, float* pos_vel,
                            mot_obstacle* obstacles) {
    if (h && !obstacles) return fail(h, MOT_ERR_INVALID, "null obstacle table");
    return ihgp_step_impl(h, rings, n_tracks, track_ids, m_state, pos_vel, obstacles);
}

static int ihgp_step_impl(mot_handle* h, const float* rings, int n_tracks, const int32_t* ids, double* m_state, float* pos_vel,
                          mot_obstacle* obstacles) {
    if (!h) return MOT_ERR_INVALID;
    if (!h->ihgp_ready) return fail(h, MOT_ERR_STATE, "mot_ihgp_configure has not been called");
    if (n_tracks < 0 || (size_t)n_tracks > h->max_tracks) return fail(h, MOT_ERR_CAPACITY, "more tracks than the handle's max_tracks");
    if (n_tracks == 0) return MOT_OK;
    if (!rings || !m_state || !pos_vel) return fail(h, MOT_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->device));
    const int L = h->ihgp_L;
    const size_t ring_elems = (size_t)n_tracks * L;
    if (ring_elems > h->ring_capacity) {
        if (h->d_rings) cudaFree(h->d_rings);
        h->d_rings = nullptr;
        CK(dalloc(&h->d_rings, h->max_tracks * (size_t)L));
        h->ring_capacity = h->max_tracks * (size_t)L;
    }
    cudaStream_t st = h->stream;
    h->prof.launches = 0;
    CK(cudaMemcpyAsync(h->d_rings, rings, ring_elems * 16, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_mstate, m_state, (size_t)n_tracks * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (obstacles && ids) CK(cudaMemcpyAsync(h->d_track_ids, ids, (size_t)n_tracks * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const int epw = ihgp_entries_per_warp(L, 112 * 1024, n_tracks, h->num_sms);
    const size_t smem = ihgp_smem_bytes(L, epw);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_ihgp_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = ((n_tracks + epw - 1) / epw + IHGP_WARPS - 1) / IHGP_WARPS;
    if (grid > h->num_sms * 8) grid = h->num_sms * 8;
    LAUNCH(KID_IHGP, k_ihgp_step<<<grid, IHGP_WARPS * 32, smem, st>>>(h->d_rings, n_tracks, L, (float)h->ihgp_dt, h->ihgp_tau, h->ihgp_axis[0],
                                                                      h->ihgp_axis[1], h->d_mstate, h->d_posvel, (obstacles && ids) ? h->d_track_ids : nullptr,
                                                                      obstacles ? h->d_obstacles : nullptr, nullptr, nullptr, 0, epw));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(m_state, h->d_mstate, (size_t)n_tracks * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pos_vel, h->d_posvel, (size_t)n_tracks * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (obstacles) CK(cudaMemcpyAsync(obstacles, h->d_obstacles, (size_t)n_tracks * sizeof(ObstacleRow), cudaMemcpyDeviceToHost, st));
    CK(mot_sync(h));
    fold_profile(h);
    return MOT_OK;
}


// ---- data association + track lifecycle on the device -----------------------------------------------------
int mot_tracks_reset(mot_handle* h) {
    if (!h) return MOT_ERR_INVALID;
    if (h->max_tracks == 0) return fail(h, MOT_ERR_CAPACITY, "handle was created with max_tracks = 0");
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_trk_meta, 0, TM_N * sizeof(int), h->stream));
    CK(mot_sync(h));
    h->trk_cur = 0; h->trk_spin = 0; h->trk_n = 0; h->trk_first = true;
    return MOT_OK;
}

// The double S with  float(sqrt(s)) < thr  <=>  s < S  for every s >= 0: both roundings are monotone, so the set of passing s is an
// initial segment of the non-negative doubles, whose bit patterns are ordered like the values (bisection over the patterns).
static double assoc_match_below(float thr) {
    if (!(thr > 0.0f)) return 0.0;  // nothing is closer than a non-positive threshold (first frame: -1)
    auto from_bits = [](uint64_t b) { double d; std::memcpy(&d, &b, sizeof(d)); return d; };
    uint64_t lo = 0, hi = 0x7ff0000000000000ull;  // s = 0 passes, s = +inf does not
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if ((float)std::sqrt(from_bits(mid)) < thr) lo = mid; else hi = mid;
    }
    return from_bits(hi);
}

double mot_assoc_match_below(float id_threshold) { return assoc_match_below(id_threshold); }

int mot_tracks_step(mot_handle* h, const float* centroids_xyzi, int n_centroids, double now, float id_threshold, float frequency,
                    int32_t* this_obj_ids, float* pos_vel, mot_obstacle* obstacles, int32_t* n_tracks, int32_t* produced) {
    if (!h) return MOT_ERR_INVALID;
    if (!h->ihgp_ready) return fail(h, MOT_ERR_STATE, "mot_ihgp_configure has not been called");
    if (h->max_tracks == 0) return fail(h, MOT_ERR_CAPACITY, "handle was created with max_tracks = 0");
    if (n_centroids < 0 || (size_t)n_centroids > h->max_tracks) return fail(h, MOT_ERR_CAPACITY, "more centroids than max_tracks");
    if (produced) *produced = 0;
    if (n_tracks) *n_tracks = h->trk_n;
    if (n_centroids == 0) return MOT_OK;  // "No obstacles around": the callback returns before any bookkeeping (MOT.cpp:170-174)
    if (!centroids_xyzi || !(frequency > 0.0f)) return fail(h, MOT_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int L = h->ihgp_L, K = n_centroids;
    if (h->trk_L != L) {  // (re)allocate the rings for this data_length; an existing table cannot survive a change of L
        for (int i = 0; i < 2; ++i) {
            if (h->d_trk_rings[i]) cudaFree(h->d_trk_rings[i]);
            h->d_trk_rings[i] = nullptr;
            CK(dalloc(&h->d_trk_rings[i], h->max_tracks * (size_t)L));
        }
        h->trk_L = L;
        int rc = mot_tracks_reset(h);
        if (rc != MOT_OK) return rc;
    }
    h->prof.launches = 0;
    const float dt_gp = 1 / frequency;  // MOT.cpp:159
    const int cur = h->trk_cur;
    CK(cudaMemcpyAsync(h->d_centroids_in, centroids_xyzi, (size_t)K * 16, cudaMemcpyDefault, st));
    // first frame: every centroid registers a track, nothing is filtered or published (MOT.cpp:126-161)
    const float thr = h->trk_first ? -1.0f : id_threshold;
    if (h->assoc_fast && h->max_tracks <= (size_t)AF_MAX_TRACKS) {
        // the sequential part against a shared-memory table of last observations, then everything a match implies in parallel
        const int cap = (int)h->max_tracks;
        const size_t smem = af_smem_bytes(cap);
        // (the kernel also has ~6 KB of static shared memory: the opt-in is needed well below 48 KB of dynamic size)
        CK(cudaFuncSetAttribute(k_associate_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 16 * 1024)));
        // CTA width by table size: the table can grow by K inside the call; the last warp keeps the records
        // (~80 instructions per warp and centroid, most of them independent of the number of tracks a thread compares: a thread takes
        // about four tracks -- 515 -> 434 us for 1,000 centroids against 1,000 tracks)
        const int af_threads = std::min(ASSOC_THREADS, std::max(128, (((h->trk_n + K) / 4 + 31) / 32) * 32 + 32));
        // fp32 screening bounds around the exact threshold (tracks.cuh): below f_lo a pair matches, above f_hi it does not
        const double mb = assoc_match_below(thr);
        float f_lo = -1.0f, f_hi = -1.0f;  // thr <= 0: every pair is "above"
        if (mb > 0.0) {
            f_hi = std::nextafterf((float)(mb * (1.0 + 2e-6)), INFINITY);
            f_lo = mb >= 1e-30 ? std::nextafterf((float)(mb * (1.0 - 2e-6)), 0.0f) : -1.0f;  // (tiny thresholds: fp32 would underflow, everything near goes to fp64)
        }
        LAUNCH(KID_ASSOCIATE, k_associate_fast<<<1, af_threads, smem, st>>>(h->d_centroids_in, K, L, cap, mb, f_lo, f_hi, h->d_trk_ids[cur], h->d_trk_rings[cur],
                                                                              h->d_trk_meta, h->d_ent_slot, h->d_ent_occ, h->d_ent_next, h->d_trk_seen,
                                                                              h->d_ent_prev, cap));
        LAUNCH(KID_TRACKS_APPLY, k_tracks_apply<<<(K + TA_THREADS / 32 - 1) / (TA_THREADS / 32), TA_THREADS, 0, st>>>(
                                     h->d_centroids_in, K, L, dt_gp, h->d_trk_ids[cur], h->d_trk_rings[cur], h->d_trk_m[cur], h->d_ent_slot, h->d_ent_occ,
                                     h->d_ent_next, h->d_trk_seen, h->d_ent_prev, h->d_ent_ids));
        CK(cudaGetLastError());  // a failed launch must not reach the host code that reads the table's counters
    } else {
        LAUNCH(KID_ASSOCIATE, k_associate<<<1, ASSOC_THREADS, 0, st>>>(h->d_centroids_in, K, L, (int)h->max_tracks, thr, dt_gp, h->d_trk_ids[cur],
                                                                      h->d_trk_rings[cur], h->d_trk_m[cur], h->d_trk_meta, h->d_trk_seen, h->d_ent_ids,
                                                                      h->d_ent_slot, h->d_ent_occ));
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_pinned + 32, h->d_trk_meta, TM_N * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(mot_sync(h));
    h->trk_n = h->h_pinned[32 + TM_NTRACKS];
    if (n_tracks) *n_tracks = h->trk_n;
    // A full track table is not fatal (the reference has no cap): the centroids that found no slot are skipped (id -1, zero
    // rows), everything else -- filtering, the callback counter, the purge that frees slots -- runs as usual.
    const int dropped = h->h_pinned[32 + TM_OVERFLOW];
    if (dropped) h->err = "track table full (max_tracks): " + std::to_string(dropped) + " centroid(s) not registered";
    if (this_obj_ids) CK(cudaMemcpyAsync(this_obj_ids, h->d_ent_ids, (size_t)K * sizeof(int), cudaMemcpyDefault, st));
    if (h->trk_first) {
        h->trk_first = false;
        CK(mot_sync(h));
        fold_profile(h);
        return dropped ? MOT_WARN_TRACKS_FULL : MOT_OK;
    }
    // callIHGP over this_objIDs (MOT.cpp:621-662); a track that was matched twice in this frame is filtered twice, in order
    const int max_occ = std::min(std::max(h->h_pinned[32 + TM_MAX_OCC], 0), K);  // (a track is matched at most K times in a frame)
    if (dropped) {
        CK(cudaMemsetAsync(h->d_posvel, 0, (size_t)K * 2 * sizeof(float4), st));
        CK(cudaMemsetAsync(h->d_obstacles, 0, (size_t)K * sizeof(ObstacleRow), st));
    }
    const int epw = ihgp_entries_per_warp(L, 112 * 1024, K, h->num_sms);
    const size_t smem = ihgp_smem_bytes(L, epw);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_ihgp_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = ((K + epw - 1) / epw + IHGP_WARPS - 1) / IHGP_WARPS;
    if (grid > h->num_sms * 8) grid = h->num_sms * 8;
    for (int r = 0; r <= max_occ; ++r)
        LAUNCH(KID_IHGP, k_ihgp_step<<<grid, IHGP_WARPS * 32, smem, st>>>(h->d_trk_rings[cur], K, L, dt_gp, h->ihgp_tau, h->ihgp_axis[0], h->ihgp_axis[1],
                                                                          h->d_trk_m[cur], h->d_posvel, h->d_ent_ids, h->d_obstacles, h->d_ent_slot,
                                                                          h->d_ent_occ, r, epw));
    if (pos_vel) CK(cudaMemcpyAsync(pos_vel, h->d_posvel, (size_t)K * 8 * sizeof(float), cudaMemcpyDefault, st));
    if (obstacles) CK(cudaMemcpyAsync(obstacles, h->d_obstacles, (size_t)K * sizeof(ObstacleRow), cudaMemcpyDefault, st));
    // unregisterOldObstacle (MOT.cpp:545-584)
    h->trk_spin += 1;
    const double period = 5;
    if (h->trk_spin > period * frequency) {
        LAUNCH(KID_TRACKS_PURGE, k_tracks_purge<<<1, ASSOC_THREADS, 0, st>>>(h->d_trk_ids[cur], h->d_trk_rings[cur], h->d_trk_m[cur], h->d_trk_ids[cur ^ 1],
                                                                            h->d_trk_rings[cur ^ 1], h->d_trk_m[cur ^ 1], L, now, period, h->d_trk_meta));
        h->trk_cur = cur ^ 1;
        h->trk_spin = 0;
        CK(cudaMemcpyAsync(h->h_pinned + 32, h->d_trk_meta, TM_N * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaGetLastError());
    CK(mot_sync(h));
    h->trk_n = h->h_pinned[32 + TM_NTRACKS];
    if (n_tracks) *n_tracks = h->trk_n;
    if (produced) *produced = 1;
    fold_profile(h);
    return dropped ? MOT_WARN_TRACKS_FULL : MOT_OK;
}

int mot_tracks_get(mot_handle* h, int32_t* ids, float* rings, double* m_state, size_t capacity, int32_t* n_tracks) {
    if (!h) return MOT_ERR_INVALID;
    if (n_tracks) *n_tracks = h->trk_n;
    if ((size_t)h->trk_n > capacity) return fail(h, MOT_ERR_CAPACITY, "track buffers too small");
    if (h->trk_n == 0) return MOT_OK;
    CK(cudaSetDevice(h->device));
    const int cur = h->trk_cur;
    if (ids) CK(cudaMemcpyAsync(ids, h->d_trk_ids[cur], (size_t)h->trk_n * sizeof(int), cudaMemcpyDefault, h->stream));
    if (rings) CK(cudaMemcpyAsync(rings, h->d_trk_rings[cur], (size_t)h->trk_n * h->trk_L * 16, cudaMemcpyDefault, h->stream));
    if (m_state) CK(cudaMemcpyAsync(m_state, h->d_trk_m[cur], (size_t)h->trk_n * 4 * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(mot_sync(h));
    return MOT_OK;
}

}  // extern "C"
