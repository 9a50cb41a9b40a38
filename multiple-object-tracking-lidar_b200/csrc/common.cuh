// common.cuh -- shared device helpers for libmot_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace mot {

// Optional per-kernel timing (mot_set_profiling): a CUDA event pair around every launch on the handle's stream.
// Off by default; bench.py switches it on for a separate pass to attribute the step time to kernels.
struct Prof {
    bool on = false;
    cudaStream_t st = nullptr;
    std::vector<cudaEvent_t> pool;
    struct Rec { int id; cudaEvent_t e0, e1; };
    std::vector<Rec> recs;
    size_t used = 0;
    int launches = 0;
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void begin(int id) {
        ++launches;
        if (!on) return;
        Rec r{id, get(), get()};
        cudaEventRecord(r.e0, st);
        recs.push_back(r);
    }
    void end() {
        if (!on) return;
        cudaEventRecord(recs.back().e1, st);
    }
};

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Work distribution used by every streaming kernel: the grid has G blocks and block b owns the contiguous
// item range [b*chunk, min(n, (b+1)*chunk)), chunk a multiple of the kernel's tile.  Contiguous chunks keep
// every compaction / scatter stable with nothing more than an exclusive sum over G per-block counters.
struct Chunking {
    int grid;
    int chunk;
};

inline Chunking make_chunking(long long n, int tile, int max_grid) {
    Chunking c;
    long long tiles = (n + tile - 1) / tile;
    if (tiles < 1) tiles = 1;
    long long g = tiles < max_grid ? tiles : max_grid;
    long long tiles_per_block = (tiles + g - 1) / g;
    c.chunk = (int)(tiles_per_block * tile);
    c.grid = (int)((n + c.chunk - 1) / c.chunk);
    if (c.grid < 1) c.grid = 1;
    return c;
}

// Order-preserving float <-> int map so float min/max can use integer atomics.
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_to_float_bits(int i) {
    int b = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    union { int i; float f; } u; u.i = b; return u.f;
#endif
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFull, v, o);
        if (lane_id() >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).  `smem` needs 33 ints.
// Returns the exclusive prefix; *total receives the block sum (same value in every thread).
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int* total) {
    int incl = warp_inclusive_scan(v);
    const int w = warp_id(), l = lane_id();
    const int nw = blockDim.x >> 5;
    __syncthreads();  // protect smem reuse across calls
    if (l == 31) smem[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = l < nw ? smem[l] : 0;
        int si = warp_inclusive_scan(s);
        smem[l] = si - s;
        if (l == 31) smem[32] = si;
    }
    __syncthreads();
    *total = smem[32];
    return smem[w] + incl - v;
}

// Sum of counts[0 .. b) for the calling block (every thread gets the result).  G is small (<= ~1200), so each
// block simply re-reduces the prefix instead of paying a separate scan launch.  `smem` needs 33 ints.
__device__ __forceinline__ int block_prefix_of(const int* __restrict__ counts, int b, int* smem) {
    int s = 0;
    for (int i = threadIdx.x; i < b; i += blockDim.x) s += counts[i];
    s = warp_sum(s);
    __syncthreads();
    if (lane_id() == 0) smem[warp_id()] = s;
    __syncthreads();
    if (warp_id() == 0) {
        int t = lane_id() < (int)(blockDim.x >> 5) ? smem[lane_id()] : 0;
        t = warp_sum(t);
        if (lane_id() == 0) smem[32] = t;
    }
    __syncthreads();
    return smem[32];
}

// Streaming 128-bit loads / stores that do not pollute L1 (every hot array is touched once per kernel).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// L2-coherent scalar access for data other CTAs mutate concurrently (union-find parents).
__device__ __forceinline__ int ld_cg(const int* p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(int* p, int v) { __stcg(p, v); }

// The reference's distance predicate (FLANN L2_Simple<float>): ((dx*dx) + dy*dy) + dz*dz, each operation
// rounded separately -- no FMA contraction, or bridge pairs within an ulp of r^2 flip the partition.
__device__ __forceinline__ float dist2_exact(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a malformed copy must never hang the GPU.  Returns false on timeout.
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    for (int i = 0; i < (1 << 22); ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
// bytes must be a multiple of 16; src/dst 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace mot
