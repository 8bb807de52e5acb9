// Weight preparation: fold the centre-tap channel substitution into every tap.
//
// Reference semantics being re-expressed: per pixel the reference subtracts the centre-tap
// terms y[kc] * W[c][kc][KH-1][KW-1], kc < c, one channel after the other
// (inf/utils/solve_mc.py:101-110) -- a Cg-step dependent chain per pixel.  With
// A0 = that strictly-lower matrix and T = (I + A0)^-1 the same update is
//     y[p] = T x[p] - sum_{q != 0} (T W_q) y[p - q]
// so the wavefront kernel only evaluates one dense product per pixel.  The adjoint solve
// uses T^T and the transposed taps.
//
// Prepared layout (floats): [dir][group][co][KDP], row = [tap t][ci]; tap 0 holds T (it
// multiplies the right-hand side x), taps t >= 1 hold -(T W_q) for q = (t / KW, t % KW).
//
// Grid (groups x layers, 2 directions, tap slabs): several layers' weights of one geometry can
// be prepared by a single launch.  Every CTA rebuilds T in shared memory (forward
// substitution, one thread per column, 4 independent partial sums: ~Cg^2/2 cycles) and then
// forms its slab of taps as small dense products out of shared memory.
#include "ifk_internal.cuh"

namespace ifk {

constexpr int kPrepThreads = 256;

__global__ void __launch_bounds__(kPrepThreads)
prepare_kernel(const float *__restrict__ weight, float *__restrict__ prepared, int C, int Cg, int Cw,
               int KH, int KW, int KD, int KDP, int taps_per_cta, int groups, size_t weight_stride,
               size_t prepared_stride)
{
    extern __shared__ float sm[];
    const int TS = Cg + 1;                 // padded row stride: column walks hit distinct banks
    float *A = sm;                         // [Cg][TS] strictly-lower centre tap A0
    float *T = A + Cg * TS;                // [Cg][TS] T0 = (I + A0)^-1 (unit lower triangular)
    float *Wq = T + Cg * TS;               // [Cg][TS] one tap, rows = weight output channel
    const int G = blockIdx.x % groups, layer = blockIdx.x / groups, dir = blockIdx.y;
    weight += (size_t)layer * weight_stride;       // batched: one weight tensor per layer
    prepared += (size_t)layer * prepared_stride;
    const int K = KH * KW;
    const size_t tap_stride = (size_t)K;                   // between input columns
    const size_t row_stride = (size_t)Cw * tap_stride;     // between output rows
    const float *wg = weight + (size_t)G * Cg * row_stride;
    const int centre = K - 1;                              // array index of shift (0,0)
    const int tid = threadIdx.x;

    for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
        const int i = e / Cg, k = e - i * Cg;
        A[i * TS + k] = k < i ? __ldg(wg + i * row_stride + k * tap_stride + centre) : 0.f;
    }
    __syncthreads();
    // column j of T0 by forward substitution: T[i][j] = [i==j] - sum_{j<=k<i} A[i][k] T[k][j]
    for (int j = tid; j < Cg; j += kPrepThreads) {
        for (int i = 0; i < j; i++) T[i * TS + j] = 0.f;
        T[j * TS + j] = 1.f;
        for (int i = j + 1; i < Cg; i++) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int k = j;
            for (; k + 3 < i; k += 4) {
                s0 = fmaf(A[i * TS + k], T[k * TS + j], s0);
                s1 = fmaf(A[i * TS + k + 1], T[(k + 1) * TS + j], s1);
                s2 = fmaf(A[i * TS + k + 2], T[(k + 2) * TS + j], s2);
                s3 = fmaf(A[i * TS + k + 3], T[(k + 3) * TS + j], s3);
            }
            for (; k < i; k++) s0 = fmaf(A[i * TS + k], T[k * TS + j], s0);
            T[i * TS + j] = -((s0 + s1) + (s2 + s3));
        }
    }
    __syncthreads();

    float *out = prepared + ((size_t)dir * C + (size_t)G * Cg) * KDP;
    const int t_begin = blockIdx.z * taps_per_cta;
    const int t_end = t_begin + taps_per_cta < K ? t_begin + taps_per_cta : K;
    for (int t = t_begin; t < t_end; t++) {
        if (t == 0) {
            // tap 0 of the prepared row is T itself (transposed for the adjoint), plus the row padding
            for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
                const int co = e / Cg, ci = e - co * Cg;
                out[(size_t)co * KDP + ci] = dir == 0 ? T[co * TS + ci] : T[ci * TS + co];
            }
            for (int e = tid; e < Cg * (KDP - KD); e += kPrepThreads) {
                const int co = e / (KDP - KD);
                out[(size_t)co * KDP + KD + (e - co * (KDP - KD))] = 0.f;
            }
            continue;
        }
        const int qh = t / KW, qw = t - qh * KW;
        const int a = (KH - 1 - qh) * KW + (KW - 1 - qw);
        __syncthreads();                       // previous tap's readers are done with Wq
        for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
            const int r = e / Cg, c = e - r * Cg;
            Wq[r * TS + c] = __ldg(wg + r * row_stride + c * tap_stride + a);
        }
        __syncthreads();
        for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
            const int co = e / Cg, ci = e - co * Cg;
            float s0 = 0.f, s1 = 0.f;
            if (dir == 0) {          // sum_{k<=co} T[co][k] * W[k][ci]
                int k = 0;
                for (; k + 1 <= co; k += 2) {
                    s0 = fmaf(T[co * TS + k], Wq[k * TS + ci], s0);
                    s1 = fmaf(T[co * TS + k + 1], Wq[(k + 1) * TS + ci], s1);
                }
                if (k <= co) s0 = fmaf(T[co * TS + k], Wq[k * TS + ci], s0);
            } else {                 // sum_{k>=co} T[k][co] * W[ci][k]
                int k = co;
                for (; k + 1 < Cg; k += 2) {
                    s0 = fmaf(T[k * TS + co], Wq[ci * TS + k], s0);
                    s1 = fmaf(T[(k + 1) * TS + co], Wq[ci * TS + k + 1], s1);
                }
                if (k < Cg) s0 = fmaf(T[k * TS + co], Wq[ci * TS + k], s0);
            }
            out[(size_t)co * KDP + t * Cg + ci] = -(s0 + s1);
        }
    }
}

int launch_prepare(const Geometry &g, const float *weight, float *prepared, cudaStream_t s, int count,
                   size_t weight_stride, size_t prepared_stride)
{
    if (count <= 0) return 0;
    const size_t smem = (size_t)3 * g.Cg * (g.Cg + 1) * sizeof(float);
    if (smem > (size_t)kMaxSmemBytes) return IFK_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    // small problems: one CTA per (group, direction) walks all taps; otherwise one tap per CTA
    // one tap per CTA unless that would mean far more CTAs than the GPU holds at once
    int taps_per_cta = 1;
    while ((long)g.groups * count * 2 * ((g.K + taps_per_cta - 1) / taps_per_cta) > 8L * kNumSM && taps_per_cta < g.K)
        taps_per_cta++;
    dim3 grid(g.groups * count, 2, (g.K + taps_per_cta - 1) / taps_per_cta);
    prepare_kernel<<<grid, kPrepThreads, smem, s>>>(weight, prepared, g.C, g.Cg, g.Cw, g.KH, g.KW, g.KD,
                                                    g.KDP, taps_per_cta, g.groups, weight_stride,
                                                    prepared_stride);
    const int st = cuda_status(cudaGetLastError());
    if (st != 0) return st;
    // the pipelined wavefront kernel reads a lane-major packed copy of these rows (ifk_solve_wave.cu)
    return launch_wave_pack(g, prepared, count, prepared_stride, s);
}

}  // namespace ifk
