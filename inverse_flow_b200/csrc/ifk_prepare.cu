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
//
// The products are what the launch costs at wide groups (Cg = 48: 864 CTAs x Cg^3 multiply-adds, each
// fed by two shared-memory loads -- the whole chip's LDS bandwidth for ~100 us, ahead of the first
// solve of the step).  They are register-tiled: a thread forms four adjacent input columns of one
// output row, per k one broadcast load of T and one 128-bit load of the staged tap (which is stored
// transposed for the adjoint, so that both directions read rows): 3 instead of 8 shared-memory
// wavefronts per 128 multiply-adds.  IFK_PREP_CFG="legacy,taps_per_cta" pins the round-1 loop / the slab size.
#include "ifk_env.cuh"
#include "ifk_internal.cuh"

namespace ifk {

constexpr int kPrepThreads = 256;

// row stride of the staged tap of the tiled product: rows 16-byte aligned, and an odd multiple of 4 so that the
// transposing stores of the adjoint spread over 8 banks
__host__ __device__ inline int prepare_tap_stride(int Cg)
{
    int rs = (Cg + 3) / 4 * 4;
    return rs % 8 == 0 ? rs + 4 : rs;
}
size_t prepare_smem_bytes(int Cg)
{
    const int ts = Cg + 1, rs = prepare_tap_stride(Cg);
    return ((size_t)2 * Cg * ts + (size_t)Cg * (rs > ts ? rs : ts)) * sizeof(float);
}

__global__ void __launch_bounds__(kPrepThreads)
prepare_kernel(const float *__restrict__ weight, float *__restrict__ prepared, int C, int Cg, int Cw,
               int KH, int KW, int KD, int KDP, int taps_per_cta, int groups, size_t weight_stride,
               size_t prepared_stride, int tiled)
{
    extern __shared__ __align__(16) float sm[];
    const int TS = Cg + 1;                 // padded row stride: column walks hit distinct banks
    const int RS = prepare_tap_stride(Cg);
    float *A = sm;                         // [Cg][TS] strictly-lower centre tap A0
    float *T = A + Cg * TS;                // [Cg][TS] T0 = (I + A0)^-1 (unit lower triangular)
    float *Wq = T + Cg * TS;               // one tap: legacy [Cg][TS], rows = weight output channel; tiled [k][RS], see below
                                           // (2 Cg (Cg+1) floats in front of it: a multiple of 16 bytes)
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
    if (tiled)                              // pad columns of the staged tap: read by the last tile, never written again
        for (int e = tid; e < Cg * (RS - Cg); e += kPrepThreads) {
            const int k = e / (RS - Cg);
            Wq[k * RS + Cg + (e - k * (RS - Cg))] = 0.f;
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
        if (tiled) {
            // out[co][ci] = -sum_k L[co][k] R[k][ci] with R[k][ci] = W[k][ci], k <= co (forward: L = T) or
            // R[k][ci] = W[ci][k], k >= co (adjoint: L = T^T, read in place with stride TS)
            for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
                const int r = e / Cg, c = e - r * Cg;
                const float w = __ldg(wg + r * row_stride + c * tap_stride + a);
                Wq[dir == 0 ? r * RS + c : c * RS + r] = w;
            }
            __syncthreads();
            const int Cg4 = (Cg + 3) >> 2;
            for (int e = tid; e < Cg * Cg4; e += kPrepThreads) {
                const int co = e / Cg4, j4 = (e - co * Cg4) * 4;
                const int k0 = dir == 0 ? 0 : co, k1 = dir == 0 ? co + 1 : Cg;
                const float *lp = dir == 0 ? T + co * TS : T + co;
                const int ls = dir == 0 ? 1 : TS;
                const float *rp = Wq + j4;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (int k = k0; k < k1; k++) {
                    const float l = lp[k * ls];
                    const float4 r4 = *reinterpret_cast<const float4 *>(rp + k * RS);
                    a0 = fmaf(l, r4.x, a0);
                    a1 = fmaf(l, r4.y, a1);
                    a2 = fmaf(l, r4.z, a2);
                    a3 = fmaf(l, r4.w, a3);
                }
                float *o = out + (size_t)co * KDP + t * Cg + j4;
                o[0] = -a0;
                if (j4 + 1 < Cg) o[1] = -a1;
                if (j4 + 2 < Cg) o[2] = -a2;
                if (j4 + 3 < Cg) o[3] = -a3;
            }
            continue;
        }
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
    const size_t smem = prepare_smem_bytes(g.Cg);
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
    const EnvKnobs &knobs = env();
    if (knobs.prep_cfg[1] > 0) taps_per_cta = knobs.prep_cfg[1] < g.K ? knobs.prep_cfg[1] : g.K;
    const int tiled = knobs.prep_cfg[0] == 1 ? 0 : 1;
    dim3 grid(g.groups * count, 2, (g.K + taps_per_cta - 1) / taps_per_cta);
    prepare_kernel<<<grid, kPrepThreads, smem, s>>>(weight, prepared, g.C, g.Cg, g.Cw, g.KH, g.KW, g.KD,
                                                    g.KDP, taps_per_cta, g.groups, weight_stride,
                                                    prepared_stride, tiled);
    const int st = cuda_status(cudaGetLastError());
    if (st != 0) return st;
    // the pipelined wavefront kernel reads a lane-major packed copy of these rows (ifk_solve_wave.cu)
    return launch_wave_pack(g, prepared, count, prepared_stride, s);
}

}  // namespace ifk
