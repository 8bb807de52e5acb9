// Issue rate of FP32 FMAs from ONE warp (and from 2 / 4 warps per scheduler) when every FMA names its own
// weight register, as the wavefront kernels do: NW weight pairs held in registers, NA accumulators,
// ND data operands reloaded from shared memory per repetition.  Prints cycles per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma_issue tools/ubench/ffma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c)
{
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int NW, int NA, bool PACKED>
__global__ void k(const float *w, float *out, long long *cyc, int reps)
{
    __shared__ float sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 1.0f + threadIdx.x * 1e-3f;
    __syncthreads();
    constexpr int ND = NW / NA;
    long long t0 = 0, t1 = 0;
    if (PACKED) {
        f32x2_t wr[NW], acc[NA];
#pragma unroll
        for (int i = 0; i < NW; i++) wr[i] = reinterpret_cast<const f32x2_t *>(w)[i * 32 + (threadIdx.x & 31)];
#pragma unroll
        for (int i = 0; i < NA; i++) acc[i] = 0ull;
        t0 = clock64();
        for (int r = 0; r < reps; r++) {
            f32x2_t d[ND];
#pragma unroll
            for (int j = 0; j < ND; j++) d[j] = reinterpret_cast<volatile f32x2_t *>(sh)[(j + r) & 31];
#pragma unroll
            for (int j = 0; j < ND; j++)
#pragma unroll
                for (int a = 0; a < NA; a++) acc[a] = fma2(wr[j * NA + a], d[j], acc[a]);
        }
        t1 = clock64();
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NA; i++) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        float wr[NW], acc[NA];
#pragma unroll
        for (int i = 0; i < NW; i++) wr[i] = w[i * 32 + (threadIdx.x & 31)];
#pragma unroll
        for (int i = 0; i < NA; i++) acc[i] = 0.f;
        t0 = clock64();
        for (int r = 0; r < reps; r++) {
            float d[ND];
#pragma unroll
            for (int j = 0; j < ND; j++) d[j] = reinterpret_cast<volatile float *>(sh)[(j + r) & 63];
#pragma unroll
            for (int j = 0; j < ND; j++)
#pragma unroll
                for (int a = 0; a < NA; a++) acc[a] = fmaf(wr[j * NA + a], d[j], acc[a]);
        }
        t1 = clock64();
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NA; i++) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NW, int NA, bool PACKED>
void run(const char *name, const float *w, float *out, long long *cyc)
{
    const int reps = 200;
    for (int warps : {1, 4, 8, 16}) {          // 4 warps = 1 per scheduler
        k<NW, NA, PACKED><<<1, warps * 32>>>(w, out, cyc, reps);
        k<NW, NA, PACKED><<<1, warps * 32>>>(w, out, cyc, reps);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / reps / NW;
        printf("%-28s warps/CTA %2d (%.2f per scheduler): %.2f cycles per instruction, %.2f FMA/lane/cycle/scheduler\n", name, warps,
               warps / 4.0, per, (PACKED ? 2.0 : 1.0) / per * (warps < 4 ? 1 : warps / 4.0));
    }
}
int main()
{
    float *w, *out;
    long long *cyc;
    cudaMalloc(&w, 256 * 32 * 8);
    cudaMemset(w, 0, 256 * 32 * 8);
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 64);
    run<66, 6, true>("FFMA2 66 weights x 6 acc", w, out, cyc);
    run<36, 6, true>("FFMA2 36 weights x 6 acc", w, out, cyc);
    run<64, 8, true>("FFMA2 64 weights x 8 acc", w, out, cyc);
    run<24, 4, true>("FFMA2 24 weights x 4 acc", w, out, cyc);
    run<132, 12, false>("FFMA 132 weights x 12 acc", w, out, cyc);
    run<72, 12, false>("FFMA 72 weights x 12 acc", w, out, cyc);
    run<48, 8, false>("FFMA 48 weights x 8 acc", w, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
