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
// Two launches per batch of layers (several layers' weights of one geometry are prepared together):
//
//  1. prepare_t_kernel, one CTA per (layer, group): T by forward substitution in shared memory (one thread
//     per column, 4 independent partial sums), written as tap 0 of BOTH directions (the adjoint's copy
//     transposed) together with the row padding.  The substitution is a Cg-step dependent chain (~8 us at
//     Cg = 48): it runs once per layer -- the first round-2 kernel repeated it in every one of the
//     2 x K CTAs of a layer, which was two thirds of the 89 us a batch of 48 Cg = 48 layers took.
//  2. prepare_taps_kernel, grid (layers x groups, 2 directions, tap slabs): reads its direction's tap 0
//     back (L = T or T^T, row-major either way), stages one tap W_q (transposed for the adjoint, so that
//     both directions read rows) and forms -L W_q register-tiled: a thread owns four adjacent input
//     columns of one output row, per k one broadcast load of L and one 128-bit load of the staged tap.
//
// IFK_PREP_CFG="0,taps_per_cta" pins the slab size of launch 2 (tests).
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
    // prepare_t_kernel: A and T, [Cg][Cg + 1] each; prepare_taps_kernel: L [Cg][Cg + 1] and the staged tap [Cg][RS]
    const size_t ts = Cg + 1, rs = prepare_tap_stride(Cg);
    const size_t a = 2 * Cg * ts, b = Cg * ts + Cg * rs + 4;
    return (a > b ? a : b) * sizeof(float);
}

struct PrepParams {
    const float *weight;
    float *prepared;
    size_t weight_stride, prepared_stride;     // floats between layers
    int C, Cg, Cw, KH, KW, KD, KDP, groups, taps_per_cta;
};

// Column j of T = (I + A0)^-1 with the column in REGISTERS (compile-time group width): row i only needs the
// thread's own earlier entries, so the shared-memory round trip per multiply-add of the generic loop (store T[i][j],
// load it back for every later row: ~30 us for 48 layers of Cg = 48) shrinks to one broadcast load of A[i][k];
// four partial sums as in the generic loop.  All lanes run the same instruction stream (entries above the diagonal
// are computed as sums of zeros and then forced to +0).
template <int CG>
__device__ __forceinline__ void t_column_in_registers(const float *A, float *T, int TS, int j)
{
    float t[CG];
#pragma unroll
    for (int i = 0; i < CG; i++) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < i; k++) s[k & 3] = fmaf(A[i * TS + k], t[k], s[k & 3]);
        const float v = -((s[0] + s[1]) + (s[2] + s[3]));
        t[i] = i < j ? 0.f : (i == j ? 1.f : v);
    }
#pragma unroll
    for (int i = 0; i < CG; i++) T[i * TS + j] = t[i];
}

__global__ void __launch_bounds__(kPrepThreads)
prepare_t_kernel(const PrepParams q)
{
    extern __shared__ __align__(16) float sm[];
    const int Cg = q.Cg, TS = Cg + 1, KDP = q.KDP, KD = q.KD;
    float *A = sm;                         // [Cg][TS] strictly-lower centre tap A0
    float *T = A + Cg * TS;                // [Cg][TS] T0 = (I + A0)^-1 (unit lower triangular)
    const int G = blockIdx.x % q.groups, layer = blockIdx.x / q.groups;
    const int K = q.KH * q.KW;
    const size_t tap_stride = (size_t)K;                   // between input columns
    const size_t row_stride = (size_t)q.Cw * tap_stride;   // between output rows
    const float *wg = q.weight + (size_t)layer * q.weight_stride + (size_t)G * Cg * row_stride;
    const int centre = K - 1;                              // array index of shift (0,0)
    const int tid = threadIdx.x;
    // programmatic dependent launch: the tap kernel behind this one may be scheduled now (its launch latency and
    // shared-memory set-up overlap the substitution); it waits for this grid before it reads T
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
        const int i = e / Cg, k = e - i * Cg;
        A[i * TS + k] = k < i ? __ldg(wg + i * row_stride + k * tap_stride + centre) : 0.f;
    }
    __syncthreads();
    // column j of T0 by forward substitution: T[i][j] = [i==j] - sum_{j<=k<i} A[i][k] T[k][j]
    const bool in_registers = Cg == 4 || Cg == 6 || Cg == 8 || Cg == 12 || Cg == 24 || Cg == 48;   // the reference models' group widths
    if (in_registers) {
        if (tid < Cg) {
            switch (Cg) {
                case 4: t_column_in_registers<4>(A, T, TS, tid); break;
                case 6: t_column_in_registers<6>(A, T, TS, tid); break;
                case 8: t_column_in_registers<8>(A, T, TS, tid); break;
                case 12: t_column_in_registers<12>(A, T, TS, tid); break;
                case 24: t_column_in_registers<24>(A, T, TS, tid); break;
                default: t_column_in_registers<48>(A, T, TS, tid); break;
            }
        }
    } else
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
    // tap 0 of the prepared rows is T itself (transposed for the adjoint), plus the row padding
    float *out = q.prepared + (size_t)layer * q.prepared_stride;
    for (int dir = 0; dir < 2; dir++) {
        float *o = out + ((size_t)dir * q.C + (size_t)G * Cg) * KDP;
        for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
            const int co = e / Cg, ci = e - co * Cg;
            o[(size_t)co * KDP + ci] = dir == 0 ? T[co * TS + ci] : T[ci * TS + co];
        }
        for (int e = tid; e < Cg * (KDP - KD); e += kPrepThreads) {
            const int co = e / (KDP - KD);
            o[(size_t)co * KDP + KD + (e - co * (KDP - KD))] = 0.f;
        }
    }
}

__global__ void __launch_bounds__(kPrepThreads)
prepare_taps_kernel(const PrepParams q)
{
    extern __shared__ __align__(16) float sm[];
    const int Cg = q.Cg, TS = Cg + 1, KDP = q.KDP, KW = q.KW, KH = q.KH;
    const int RS = prepare_tap_stride(Cg);
    float *L = sm;                                         // [Cg][TS]: T (forward) or T^T (adjoint)
    float *Wq = L + (Cg * TS + 3) / 4 * 4;                 // [k][RS] one tap, rows 16-byte aligned
    const int G = blockIdx.x % q.groups, layer = blockIdx.x / q.groups, dir = blockIdx.y;
    const int K = KH * KW;
    const size_t tap_stride = (size_t)K, row_stride = (size_t)q.Cw * tap_stride;
    const float *wg = q.weight + (size_t)layer * q.weight_stride + (size_t)G * Cg * row_stride;
    float *out = q.prepared + (size_t)layer * q.prepared_stride + ((size_t)dir * q.C + (size_t)G * Cg) * KDP;
    const int tid = threadIdx.x;

    for (int e = tid; e < Cg * (RS - Cg); e += kPrepThreads) {   // pad columns: read by the last tile, never written again
        const int k = e / (RS - Cg);
        Wq[k * RS + Cg + (e - k * (RS - Cg))] = 0.f;
    }
    // T comes from the launch before (prepare_t_kernel); dependents (the pack kernel) are released after the wait, so
    // that "everything before my predecessor is complete" holds for them too
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
        const int co = e / Cg, k = e - co * Cg;
        L[co * TS + k] = out[(size_t)co * KDP + k];
    }
    const int t_begin = 1 + blockIdx.z * q.taps_per_cta;
    const int t_end = t_begin + q.taps_per_cta < K ? t_begin + q.taps_per_cta : K;
    const int Cg4 = (Cg + 3) >> 2;
    for (int t = t_begin; t < t_end; t++) {
        const int qh = t / KW, qw = t - qh * KW;
        const int a = (KH - 1 - qh) * KW + (KW - 1 - qw);
        __syncthreads();                       // previous tap's readers are done with Wq
        // out[co][ci] = -sum_k L[co][k] R[k][ci] with R[k][ci] = W[k][ci], k <= co (forward) or R[k][ci] = W[ci][k], k >= co (adjoint)
        for (int e = tid; e < Cg * Cg; e += kPrepThreads) {
            const int r = e / Cg, c = e - r * Cg;
            const float w = __ldg(wg + r * row_stride + c * tap_stride + a);
            Wq[dir == 0 ? r * RS + c : c * RS + r] = w;
        }
        __syncthreads();
        for (int e = tid; e < Cg * Cg4; e += kPrepThreads) {
            const int co = e / Cg4, j4 = (e - co * Cg4) * 4;
            const int k0 = dir == 0 ? 0 : co, k1 = dir == 0 ? co + 1 : Cg;
            const float *lp = L + co * TS;
            const float *rp = Wq + j4;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int k = k0;
            for (; k + 3 < k1; k += 4) {               // the eight loads of four steps in flight together
                float l[4];
                float4 r4[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    l[u] = lp[k + u];
                    r4[u] = *reinterpret_cast<const float4 *>(rp + (k + u) * RS);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    a0 = fmaf(l[u], r4[u].x, a0);
                    a1 = fmaf(l[u], r4[u].y, a1);
                    a2 = fmaf(l[u], r4[u].z, a2);
                    a3 = fmaf(l[u], r4[u].w, a3);
                }
            }
            for (; k < k1; k++) {
                const float l = lp[k];
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
    }
}

int launch_prepare(const Geometry &g, const float *weight, float *prepared, cudaStream_t s, int count,
                   size_t weight_stride, size_t prepared_stride)
{
    if (count <= 0) return 0;
    const size_t smem = prepare_smem_bytes(g.Cg);
    if (smem > (size_t)kMaxSmemBytes) return IFK_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(prepare_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(prepare_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    PrepParams q{};
    q.weight = weight; q.prepared = prepared;
    q.weight_stride = weight_stride; q.prepared_stride = prepared_stride;
    q.C = g.C; q.Cg = g.Cg; q.Cw = g.Cw; q.KH = g.KH; q.KW = g.KW; q.KD = g.KD; q.KDP = g.KDP; q.groups = g.groups;
    prepare_t_kernel<<<g.groups * count, kPrepThreads, smem, s>>>(q);
    int st = cuda_status(cudaGetLastError());
    if (st != 0) return st;
    if (g.K > 1) {
        // one tap per CTA unless that would mean far more CTAs than the GPU holds at once
        int taps_per_cta = 1;
        while ((long)g.groups * count * 2 * ((g.K - 1 + taps_per_cta - 1) / taps_per_cta) > 8L * kNumSM && taps_per_cta < g.K - 1)
            taps_per_cta++;
        const EnvKnobs &knobs = env();
        if (knobs.prep_cfg[1] > 0) taps_per_cta = knobs.prep_cfg[1] < g.K - 1 ? knobs.prep_cfg[1] : g.K - 1;
        q.taps_per_cta = taps_per_cta;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(g.groups * count, 2, (g.K - 1 + taps_per_cta - 1) / taps_per_cta);
        cfg.blockDim = dim3(kPrepThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = env().pdl ? 1 : 0;
        st = cuda_status(cudaLaunchKernelEx(&cfg, prepare_taps_kernel, q));
        if (st != 0) return st;
    }
    // the pipelined wavefront kernel reads a lane-major packed copy of these rows (ifk_solve_wave.cu)
    return launch_wave_pack(g, prepared, count, prepared_stride, s);
}

}  // namespace ifk
