// Data-parallel sum of the dW bucket over peer memory: ONE kernel that synchronises with the other
// ranks, reads every rank's bucket through NVLink/NVSwitch peer mappings and writes the sum -- the
// only exchange on this path (the reference uses single-process nn.DataParallel,
// inf/if_multiGPU_imagenet32.py:410-411).
//
// Why not NCCL: the bucket is 20 KB (if_glow_mnist) to 5 MB (if_multiGPU_imagenet32) and the step it
// follows takes 0.2 - 3 ms; a ring/tree all-reduce of that size is latency bound (tens of
// microseconds at 8 ranks, issued from the host between two graph replays), while one-shot peer
// loads cost one NVLink round trip plus the payload, sit inside the step's CUDA graph, and sum in a
// FIXED rank order: every rank computes bit-identical gradients (a ring's order depends on the rank).
//
// Protocol (every rank launches the same grid, once per step, in the same order):
//   block b, thread r < world:   bump the block's epoch; tell rank r "my bucket is final" by writing the
//                                epoch into r's flag cell [phase 0][b][my rank]; spin until all cells
//                                [phase 0][b][*] of MY flag block have reached the epoch;
//   all threads:                 out[i] = sum_{r = 0..world-1} bucket_r[i]  (128-bit peer loads, grid-stride);
//   block b, thread r < world:   the same hand-shake on phase 1 = "I have read your bucket": nobody may
//                                overwrite its bucket (the next step's stage-2 reduce) before that.
// The buckets were written by an earlier kernel of the same stream, so they are complete when this kernel
// starts; peer loads are served by the owner's L2 and are not cached in the reader's L2.
#include <stdint.h>
#include "ifk_internal.cuh"

namespace ifk {

constexpr int kCommMaxWorld = 16;
constexpr int kCommMaxBlocks = 64;

struct CommParams {
    const float *bucket[kCommMaxWorld];     // rank r's bucket, mapped into this process
    unsigned *flags[kCommMaxWorld];         // rank r's flag block: [2][kCommMaxBlocks][kCommMaxWorld] cells + [kCommMaxBlocks] epochs
    float *out;
    size_t n;
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float *p)
{
    float4 v;       // relaxed system-scope load: never served from a stale local line
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer_f1(const float *p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void rank_barrier(const CommParams &p, int phase, unsigned epoch)
{
    const int b = blockIdx.x, t = threadIdx.x;
    __syncthreads();                                     // this block's reads / writes are issued
    if (t < p.world) {
        __threadfence_system();
        unsigned *cell = p.flags[t] + ((size_t)phase * kCommMaxBlocks + b) * kCommMaxWorld + p.rank;
        st_release_sys(cell, epoch);
        const unsigned *mine = p.flags[p.rank] + ((size_t)phase * kCommMaxBlocks + b) * kCommMaxWorld + t;
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) { }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) allreduce_peer_kernel(const CommParams p)
{
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) {
        unsigned *ep = p.flags[p.rank] + (size_t)2 * kCommMaxBlocks * kCommMaxWorld + blockIdx.x;
        s_epoch = *ep + 1;                               // only this block ever touches its epoch
        *ep = s_epoch;
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    rank_barrier(p, 0, epoch);                           // every rank's bucket is final

    const size_t n4 = p.n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = ld_peer_f4(p.bucket[0] + 4 * i);
        for (int r = 1; r < p.world; r++) {              // fixed rank order: bit-identical on every rank
            const float4 v = ld_peer_f4(p.bucket[r] + 4 * i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4 *>(p.out + 4 * i) = acc;
    }
    if (blockIdx.x == 0)
        for (size_t i = 4 * n4 + threadIdx.x; i < p.n; i += blockDim.x) {
            float acc = ld_peer_f1(p.bucket[0] + i);
            for (int r = 1; r < p.world; r++) acc += ld_peer_f1(p.bucket[r] + i);
            p.out[i] = acc;
        }
    rank_barrier(p, 1, epoch);                           // every rank has read every bucket
}

}  // namespace ifk

using namespace ifk;

extern "C" {

size_t ifk_allreduce_flag_bytes(void)
{
    return ((size_t)2 * kCommMaxBlocks * kCommMaxWorld + kCommMaxBlocks) * sizeof(unsigned);
}

int ifk_allreduce_peer_f32(const float *const *buckets, unsigned *const *flags, int rank, int world, float *out,
                           size_t n, ifk_stream_t stream)
{
    if (!buckets || !flags || !out) return IFK_ERR_NULL_POINTER;
    if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world) return IFK_ERR_BAD_SHAPE;
    CommParams p{};
    for (int r = 0; r < world; r++) {
        if (!buckets[r] || !flags[r]) return IFK_ERR_NULL_POINTER;
        if ((uintptr_t)buckets[r] % 16 != 0) return IFK_ERR_BAD_LAYOUT;
        p.bucket[r] = buckets[r];
        p.flags[r] = flags[r];
    }
    if ((uintptr_t)out % 16 != 0 || out == buckets[rank]) return IFK_ERR_BAD_LAYOUT;
    p.out = out; p.n = n; p.rank = rank; p.world = world;
    if (n == 0) return 0;
    // enough blocks to keep ~16 KB of peer loads in flight each, few enough that all are resident at once
    size_t blocks = (n / 4 + 1023) / 1024;
    if (blocks < 1) blocks = 1;
    if (blocks > (size_t)kCommMaxBlocks) blocks = kCommMaxBlocks;
    allreduce_peer_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return cuda_status(cudaGetLastError());
}

}  // extern "C"
