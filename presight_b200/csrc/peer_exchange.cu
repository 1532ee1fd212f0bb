// Gradient exchange over peer memory (NVLink / NVSwitch) without collective kernels: the building blocks of
// presight_b200/peer_exchange.py, the data-parallel exchange step of SURVEY §8e (reference: DDP's bucketed all-reduce,
// pipelines/PreSight/my_pipeline.py:121-124).
//
// Why not NCCL alone: the all-reduce of the 512 MiB main-table gradient has to travel while the atomics-bound scatter and
// the persistent proposal kernels own every SM.  An NCCL CTA (512-640 threads x ~96 registers) needs an almost empty SM
// and starves until those kernels retire; with fewer / smaller NCCL CTAs the transfer itself becomes the long pole
// (profiles/r2_e_n2_variants.txt).  Here the bytes move on the COPY ENGINES — cudaMemcpyAsync between IPC-mapped buffers of
// the ranks of one node — and the only kernels are a one-warp flag wait and the sum of the received pieces:
//   reduce-scatter  every rank pushes, for each peer, the peer's shard of its gradient into the peer's staging area, then a
//                   4-byte flag;
//   reduce          the owner waits for the flags, sums its own shard and the N-1 staged ones (ps_peer_reduce);
//   all-gather      the owner pushes the reduced shard into every peer's gradient buffer, then a flag.
// Buffers that peers write into are plain cudaMalloc allocations (ps_peer_alloc) exported with cudaIpcGetMemHandle.
#include "common.cuh"

namespace ps {

// all `n` flags (stride in elements) equal `value`: one warp polls, lane i the flags i, i + 32, ...
__global__ void peer_wait_flags_kernel(const volatile uint32_t* flags, int n, int stride, uint32_t value) {
    for (int i = threadIdx.x; i < n; i += 32) {
        // (step numbers only grow: a flag that has already moved on also satisfies the wait)
        while ((int32_t)(flags[(size_t)i * stride] - value) < 0) __nanosleep(200);
    }
    __threadfence_system();
}

// dst[i] = scale * (dst[i] + sum_k src_k[i]), 16-byte vectors, n a multiple of 4
struct PeerSrcs {
    const float4* p[PS_MAX_FIELDS];
    int n;
};
__global__ void __launch_bounds__(256) peer_reduce_kernel(float4* __restrict__ dst, PeerSrcs s, int64_t n4, float scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = dst[i];
        for (int k = 0; k < s.n; ++k) {
            const float4 b = __ldcg(s.p[k] + i);        // written by a peer's copy engine: read through L2
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        dst[i] = a;
    }
}

// ---- the same exchange with P2P stores from (few, small) CTAs instead of copy-engine transfers: one kernel pushes, one
// reduces and broadcasts; no per-copy DMA latency (at 8 ranks a range is 14 data + 16 flag copies per rank) --------------------
struct PeerRange {
    char* base[8];          // every rank's allocation (this rank's own included), as mapped in this process
    int world, rank;
    size_t g_off, s_off;    // gradient buffer / staging area inside an allocation
    size_t range_off;       // first byte of the row range inside the gradient buffer
    size_t shard_bytes;     // bytes of one rank's shard of the range (multiple of 16)
    size_t flag1_off, flag2_off;   // the range's phase flags inside an allocation: [world] uint32 each
    uint32_t step;
    float scale;
    unsigned int* counter;  // CTAs that have finished (returns to zero)
};

__device__ __forceinline__ void peer_signal(const PeerRange& a, size_t flag_off) {
    // all stores of this CTA are ordered before the count, the last CTA's flag stores behind everybody's count
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.counter, 1u);
        if (prev == gridDim.x - 1) {
            *a.counter = 0u;
            __threadfence_system();
            for (int r = 0; r < a.world; ++r)
                *reinterpret_cast<volatile uint32_t*>(a.base[r] + flag_off + 4 * a.rank) = a.step;
        }
    }
}

// my rows of every peer's shard -> that peer's staging slot `rank`
__global__ void __launch_bounds__(256) peer_push_kernel(PeerRange a) {
    const int64_t n16 = (int64_t)(a.shard_bytes / 16);
    const int64_t total = n16 * (a.world - 1);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int d = (int)(i / n16) + 1;
        const int64_t j = i - (int64_t)(d - 1) * n16;
        const int r = (a.rank + d) % a.world;
        const float4 v = *(reinterpret_cast<const float4*>(a.base[a.rank] + a.g_off + a.range_off + (size_t)r * a.shard_bytes) + j);
        *(reinterpret_cast<float4*>(a.base[r] + a.s_off + a.range_off + (size_t)a.rank * a.shard_bytes) + j) = v;
    }
    peer_signal(a, a.flag1_off);
}

// wait for every rank's pushes, average my shard, write it into every rank's gradient buffer
__global__ void __launch_bounds__(256) peer_reduce_bcast_kernel(PeerRange a) {
    if (threadIdx.x < a.world) {
        const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(a.base[a.rank] + a.flag1_off) + threadIdx.x;
        while ((int32_t)(*f - a.step) < 0) __nanosleep(200);
        __threadfence_system();
    }
    __syncthreads();
    const int64_t n16 = (int64_t)(a.shard_bytes / 16);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const size_t mine = a.range_off + (size_t)a.rank * a.shard_bytes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        float4 v = *(reinterpret_cast<const float4*>(a.base[a.rank] + a.g_off + mine) + i);
        for (int q = 0; q < a.world; ++q) {
            if (q == a.rank) continue;
            const float4 b = __ldcg(reinterpret_cast<const float4*>(a.base[a.rank] + a.s_off + a.range_off + (size_t)q * a.shard_bytes) + i);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
        for (int r = 0; r < a.world; ++r) *(reinterpret_cast<float4*>(a.base[r] + a.g_off + mine) + i) = v;
    }
    peer_signal(a, a.flag2_off);
}

}  // namespace ps

using namespace ps;

/* One row range of the exchange with kernels: push (phase 1) then reduce + broadcast (phase 2), `ctas` CTAs of 256 threads
 * each.  peer_bases_host[world]: every rank's allocation as mapped here; counters: two zero-initialised uint32 in local
 * device memory, private to this range. */
extern "C" int ps_peer_exchange_range(void* const* peer_bases_host, int world, int rank, size_t g_off, size_t s_off,
                                      size_t range_off, size_t shard_bytes, size_t flag1_off, size_t flag2_off, uint32_t step,
                                      float scale, unsigned int* counters, int ctas, void* stream) {
    PS_REQUIRE(peer_bases_host && counters, "peer_exchange_range: null pointer");
    PS_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world, "peer_exchange_range: world %d rank %d", world, rank);
    PS_REQUIRE(shard_bytes % 16 == 0 && range_off % 16 == 0 && g_off % 16 == 0 && s_off % 16 == 0,
               "peer_exchange_range: offsets must be multiples of 16 bytes");
    PS_REQUIRE(ctas >= 1 && ctas <= 1024, "peer_exchange_range: %d CTAs", ctas);
    PeerRange a{};
    for (int r = 0; r < world; ++r) {
        PS_REQUIRE(peer_bases_host[r] != nullptr, "peer_exchange_range: rank %d not mapped", r);
        a.base[r] = static_cast<char*>(peer_bases_host[r]);
    }
    a.world = world; a.rank = rank; a.g_off = g_off; a.s_off = s_off; a.range_off = range_off; a.shard_bytes = shard_bytes;
    a.flag1_off = flag1_off; a.flag2_off = flag2_off; a.step = step; a.scale = scale;
    a.counter = counters;
    peer_push_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(a);
    if (int e = check_launch("peer_push")) return e;
    a.counter = counters + 1;
    peer_reduce_bcast_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("peer_reduce_bcast");
}

extern "C" int ps_peer_alloc(size_t bytes, void** ptr) {
    PS_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: bad arguments");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) {
        set_error("peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" int ps_peer_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        set_error("peer_free: %s", cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" int ps_peer_export(void* ptr, void* handle64_host) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PS_REQUIRE(ptr && handle64_host, "peer_export: null pointer");
    cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64_host), ptr);
    if (e != cudaSuccess) {
        set_error("peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" int ps_peer_open(const void* handle64_host, void** ptr) {
    PS_REQUIRE(ptr && handle64_host, "peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" int ps_peer_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) {
        set_error("peer_close: %s", cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" int ps_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return 0;
    PS_REQUIRE(dst && src, "peer_copy: null pointer");
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("peer_copy: %s", cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" int ps_peer_wait_flags(const uint32_t* flags, int n, int stride, uint32_t value, void* stream) {
    if (n <= 0) return 0;
    PS_REQUIRE(flags != nullptr && stride >= 1, "peer_wait_flags: bad arguments");
    peer_wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n, stride, value);
    return check_launch("peer_wait_flags");
}

extern "C" int ps_peer_reduce(float* dst, const float* const* srcs_host, int n_src, int64_t n, float scale, void* stream) {
    if (n == 0) return 0;
    PS_REQUIRE(dst && (n_src == 0 || srcs_host), "peer_reduce: null pointer");
    PS_REQUIRE(n_src >= 0 && n_src <= PS_MAX_FIELDS, "peer_reduce: %d sources (max %d)", n_src, PS_MAX_FIELDS);
    PS_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "peer_reduce: n and dst must be 16-byte multiples");
    PeerSrcs s{};
    s.n = n_src;
    for (int k = 0; k < n_src; ++k) {
        PS_REQUIRE(srcs_host[k] && (reinterpret_cast<uintptr_t>(srcs_host[k]) & 15) == 0, "peer_reduce: source %d unaligned", k);
        s.p[k] = reinterpret_cast<const float4*>(srcs_host[k]);
    }
    const int64_t n4 = n / 4;
    // a few CTAs only: the sum is a small side job next to the kernels it runs beside
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    peer_reduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(dst), s, n4, scale);
    return check_launch("peer_reduce");
}
