// Measuring aids behind ifk_debug_fp32_peak / ifk_debug_latencies (ifk.h): the denominators of the
// roofline bench.py reports are MEASURED on the box, not quoted -- the FP32 FMA rate of the CUDA cores
// (scalar FFMA and packed FFMA2) and the latencies of the instructions a wavefront step chains
// (FFMA, FFMA2, warp shuffle, shared-memory load, store->barrier->load round trip).
// Not part of the product path: these calls synchronise the device.
#include <stdint.h>
#include "ifk_env.cuh"
#include "ifk_internal.cuh"
#include "ifk_solve_kernel.cuh"

namespace ifk {

// ---- throughput: NACC independent accumulators per thread, ITER rounds ------------------------------
template <int PACKED, int NACC>
__global__ void __launch_bounds__(256) fma_rate_kernel(float *out, int iters, float seed)
{
    if constexpr (PACKED) {
        f32x2_t acc[NACC], a[NACC];
        const f32x2_t b = pack_f32x2(seed, seed * 0.5f);
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            acc[i] = pack_f32x2((float)threadIdx.x * 1e-3f + i, 1.f);
            a[i] = pack_f32x2(1.0f + 1e-6f * i, 1.0f - 1e-6f * i);
        }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fma_f32x2(a[i], acc[i], b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NACC; i++) s += sum_f32x2(acc[i]);
        if (s == 12345.678f) out[0] = s;       // never true: keeps the chain alive
    } else {
        float acc[NACC], a[NACC];
        const float b = seed;
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            acc[i] = (float)threadIdx.x * 1e-3f + i;
            a[i] = 1.0f + 1e-6f * i;
        }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fmaf(a[i], acc[i], b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NACC; i++) s += acc[i];
        if (s == 12345.678f) out[0] = s;
    }
}

template <int PACKED>
static double fma_rate(int ctas_per_sm, int iters, float *scratch, cudaStream_t s)
{
    constexpr int NACC = 16;
    const int grid = device_sm_count() * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fma_rate_kernel<PACKED, NACC><<<grid, 256, 0, s>>>(scratch, iters, 1e-7f);     // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, s);
        fma_rate_kernel<PACKED, NACC><<<grid, 256, 0, s>>>(scratch, iters, 1e-7f);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * (PACKED ? 2 : 1) * NACC * (double)iters * 256.0 * grid;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

// ---- latencies: one warp (or 8), N dependent operations between two clock64() reads ------------------
// out[0] FFMA  out[1] FFMA2  out[2] SHFL.BFLY  out[3] LDS (pointer chase)  out[4] st.shared -> __syncwarp -> ld.shared
// out[5] st.shared -> bar.sync (8 warps) -> ld.shared   out[6] bar.sync alone (8 warps)   out[7] FADD
__global__ void __launch_bounds__(256) latency_kernel(long long *out, int n)
{
    __shared__ int chase[64];
    __shared__ float cell[256];
    const int tid = threadIdx.x;
    if (tid < 64) chase[tid] = (tid + 1) & 63;
    cell[tid] = 0.f;
    __syncthreads();
    float x = (float)tid * 1e-3f, a = 1.000001f, b = 1e-7f;
    long long t0, t1;
    // FFMA
    t0 = clock64();
    for (int i = 0; i < n; i++) x = fmaf(a, x, b);
    t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    // FFMA2
    f32x2_t p2 = pack_f32x2(x, x + 1.f);
    const f32x2_t a2 = pack_f32x2(a, a), b2 = pack_f32x2(b, b);
    t0 = clock64();
    for (int i = 0; i < n; i++) p2 = fma_f32x2(a2, p2, b2);
    t1 = clock64();
    x += sum_f32x2(p2);
    if (tid == 0) out[1] = t1 - t0;
    // shuffle
    t0 = clock64();
    for (int i = 0; i < n; i++) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64();
    if (tid == 0) out[2] = t1 - t0;
    // dependent shared-memory loads
    int idx = tid & 63;
    t0 = clock64();
    for (int i = 0; i < n; i++) idx = chase[idx];
    t1 = clock64();
    if (tid == 0) out[3] = t1 - t0;
    x += (float)idx;
    // store -> __syncwarp -> load from the neighbouring lane's cell (one warp's producer/consumer hand-off)
    t0 = clock64();
    for (int i = 0; i < n; i++) {
        cell[tid] = x;
        __syncwarp();
        x = cell[tid ^ 1] + 1.f;
        __syncwarp();
    }
    t1 = clock64();
    if (tid == 0) out[4] = t1 - t0;
    // store -> block barrier -> load from another warp's cell (the wavefront step's hand-off)
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < n; i++) {
        cell[tid] = x;
        __syncthreads();
        x = cell[(tid + 32) & 255] + 1.f;
    }
    t1 = clock64();
    if (tid == 0) out[5] = t1 - t0;
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < n; i++) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[6] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < n; i++) x = x + b;
    t1 = clock64();
    if (tid == 0) out[7] = t1 - t0;
    if (x == 12345.678f) out[8] = (long long)x;
}

}  // namespace ifk

using namespace ifk;

extern "C" {

// TFLOP/s of dependent-chain-free FP32 FMAs on all SMs: out[0] scalar FFMA, out[1] packed FFMA2
// (2 flops per FMA).  `scratch`: >= 4 bytes of device memory.  Synchronises.
int ifk_debug_fp32_peak(float *scratch, double *out_tflops)
{
    if (!scratch || !out_tflops) return IFK_ERR_NULL_POINTER;
    out_tflops[0] = fma_rate<0>(8, 1 << 14, scratch, 0);
    out_tflops[1] = fma_rate<1>(8, 1 << 14, scratch, 0);
    return cuda_status(cudaDeviceSynchronize());
}

// cycles per dependent operation (see latency_kernel); `device_out`: 16 x int64 of device memory,
// `host_out`: 8 doubles.  Synchronises.
int ifk_debug_latencies(long long *device_out, double *host_out)
{
    if (!device_out || !host_out) return IFK_ERR_NULL_POINTER;
    const int n = 2048;
    latency_kernel<<<1, 256>>>(device_out, n);      // warm-up (instruction cache)
    latency_kernel<<<1, 256>>>(device_out, n);
    long long h[8];
    cudaError_t e = cudaMemcpy(h, device_out, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < 8; i++) host_out[i] = (double)h[i] / n;
    return 0;
}

}  // extern "C"
