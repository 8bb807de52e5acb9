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
#include "ifk_internal.cuh"

namespace ifk {

__global__ void __launch_bounds__(256)
prepare_kernel(const float *__restrict__ weight, float *__restrict__ prepared, int C, int Cg,
               int Cw, int KH, int KW, int KD, int KDP)
{
    extern __shared__ float T[];  // T0 = (I + A0)^-1, row-major Cg x Cg (unit lower triangular)
    const int G = blockIdx.x, dir = blockIdx.y;
    const int K = KH * KW;
    const size_t tap_stride = (size_t)KH * KW;             // between input columns
    const size_t row_stride = (size_t)Cw * tap_stride;     // between output rows
    const float *wg = weight + (size_t)G * Cg * row_stride;
    const int centre = K - 1;                              // array index of shift (0,0)

    // column j of T0 by forward substitution, one thread per column, double accumulation
    for (int j = threadIdx.x; j < Cg; j += blockDim.x) {
        for (int i = 0; i < Cg; i++) {
            double acc = (i == j) ? 1.0 : 0.0;
            if (i > j) {
                for (int k = j; k < i; k++)
                    acc -= (double)wg[i * row_stride + k * tap_stride + centre] * (double)T[k * Cg + j];
            } else if (i < j) {
                acc = 0.0;
            }
            T[i * Cg + j] = (float)acc;
        }
    }
    __syncthreads();

    float *out = prepared + ((size_t)dir * C + (size_t)G * Cg) * KDP;
    const int total = Cg * KDP;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int co = e / KDP, kidx = e - co * KDP;
        float v = 0.f;
        if (kidx < KD) {
            const int t = kidx / Cg, ci = kidx - t * Cg;
            if (t == 0) {
                v = dir == 0 ? T[co * Cg + ci] : T[ci * Cg + co];
            } else {
                const int qh = t / KW, qw = t - qh * KW;
                const int a = (KH - 1 - qh) * KW + (KW - 1 - qw);
                double acc = 0.0;
                if (dir == 0) {      // sum_k T0[co][k] * W[k][ci][a]
                    for (int k = 0; k <= co; k++)
                        acc += (double)T[co * Cg + k] * (double)wg[k * row_stride + ci * tap_stride + a];
                } else {             // sum_k T0[k][co] * W[ci][k][a]
                    for (int k = co; k < Cg; k++)
                        acc += (double)T[k * Cg + co] * (double)wg[ci * row_stride + k * tap_stride + a];
                }
                v = (float)(-acc);
            }
        }
        out[e] = v;
    }
}

int launch_prepare(const Geometry &g, const float *weight, float *prepared, cudaStream_t s)
{
    const size_t smem = (size_t)g.Cg * g.Cg * sizeof(float);
    if (smem > (size_t)kMaxSmemBytes) return IFK_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(prepare_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(g.groups, 2);
    prepare_kernel<<<grid, 256, smem, s>>>(weight, prepared, g.C, g.Cg, g.Cw, g.KH, g.KW, g.KD, g.KDP);
    return cuda_status(cudaGetLastError());
}

}  // namespace ifk
