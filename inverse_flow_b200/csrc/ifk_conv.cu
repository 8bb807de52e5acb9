// x = L y : the masked (autoregressive) convolution, sampling direction z -> x.
// Replaces inv_conv_fwd_cuda_inverse (inv_conv_with_bp_kernel_general.cu:141-264), which
// walks the anti-diagonals with a launch + device sync each although nothing depends on
// anything here.
//
// conv_tiled_kernel: a CTA owns a stripe of images of one channel group.  The group's masked
// weights are transposed once into shared memory as [tap][ci][co] (centre tap = unit diagonal +
// strictly-lower triangle, so the mask costs nothing afterwards); a thread owns (pixel, 4
// output channels) items: per (tap, ci) one coalesced global load of y (neighbouring pixels
// overlap in L1) and one broadcast 128-bit shared load of 4 weights feed 4 FMAs; the border
// test is per tap, not per element.
// conv_wide_kernel: wide groups (Cg > 16) whose staged weights fit shared memory: 4 pixels x 4 output
// channels per thread with a sliding row window (see the kernel).
// conv_kernel: one thread per output, everything through L1/L2 -- single-channel groups and whatever the
// staged kernels cannot hold.
#include <stdlib.h>
#include "ifk_env.cuh"
#include "ifk_internal.cuh"

namespace ifk {

__global__ void __launch_bounds__(256)
conv_tiled_kernel(const float *__restrict__ y, const float *__restrict__ weight, float *__restrict__ x,
                  int B, int C, int H, int W, int KH, int KW, int Cw, int Cg, int CgP4, int nsplit, int orient)
{
    extern __shared__ __align__(16) float wT[];              // [K][Cg][CgP4]
    const int HW = H * W, K = KH * KW;
    const int G = blockIdx.y, tid = threadIdx.x;
    const size_t tap_stride = (size_t)K, row_stride = (size_t)Cw * tap_stride;
    const float *wg = weight + (size_t)G * Cg * row_stride;

    for (int e = tid; e < K * Cg * CgP4; e += blockDim.x) {
        const int co = e % CgP4, ci = (e / CgP4) % Cg, t = e / (CgP4 * Cg);
        const int qh = t / KW, qw = t - qh * KW;
        const int a = (KH - 1 - qh) * KW + (KW - 1 - qw);
        float v = 0.f;
        if (co < Cg) {
            if (t == 0) v = ci < co ? __ldg(wg + co * row_stride + ci * tap_stride + a) : (ci == co ? 1.f : 0.f);
            else        v = __ldg(wg + co * row_stride + ci * tap_stride + a);
        }
        wT[e] = v;
    }
    __syncthreads();

    const int nquad = CgP4 >> 2;
    const bool fw = orient & 1, fh = orient & 2;              // reflected axes (ifk.h IFK_ORIENT_*)
    const int step_h = fh ? -W : W, step_w = fw ? -1 : 1;     // memory step towards the neighbour's side
    // work unit = (image, split): the items of one image may be shared by `nsplit` CTAs
    for (int u = blockIdx.x; u < B * nsplit; u += gridDim.x) {
        const int b = u / nsplit, sp = u - b * nsplit;
        const float *yb = y + ((size_t)b * C + (size_t)G * Cg) * HW;
        float *xb = x + ((size_t)b * C + (size_t)G * Cg) * HW;
        for (int item = sp * blockDim.x + tid; item < HW * nquad; item += nsplit * blockDim.x) {
            const int cq = item / HW, r = item - cq * HW;
            const int hm = r / W, wm = r - hm * W;
            const int h = fh ? H - 1 - hm : hm, w = fw ? W - 1 - wm : wm;     // causal-frame coordinates
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const int qh_max = h < KH - 1 ? h : KH - 1, qw_max = w < KW - 1 ? w : KW - 1;
            for (int qh = 0; qh <= qh_max; qh++)
                for (int qw = 0; qw <= qw_max; qw++) {
                    const float *yp = yb + r - qh * step_h - qw * step_w;
                    const float *wp = wT + (size_t)(qh * KW + qw) * Cg * CgP4 + cq * 4;
#pragma unroll 4
                    for (int ci = 0; ci < Cg; ci++) {
                        const float yv = __ldg(yp + (size_t)ci * HW);
                        const float4 w4 = *reinterpret_cast<const float4 *>(wp + ci * CgP4);
                        a0 = fmaf(w4.x, yv, a0);
                        a1 = fmaf(w4.y, yv, a1);
                        a2 = fmaf(w4.z, yv, a2);
                        a3 = fmaf(w4.w, yv, a3);
                    }
                }
            const int co = cq * 4;
            xb[(size_t)co * HW + r] = a0;
            if (co + 1 < Cg) xb[(size_t)(co + 1) * HW + r] = a1;
            if (co + 2 < Cg) xb[(size_t)(co + 2) * HW + r] = a2;
            if (co + 3 < Cg) xb[(size_t)(co + 3) * HW + r] = a3;
        }
    }
}

// conv_wide_kernel: wide groups (Cg > 16).  Same staged weights as conv_tiled_kernel; a thread owns a
// 4-pixel run of one image row x 4 output channels and, per (row offset qh, input channel), loads the
// KW + 3 values y(h - qh, w0 - KW + 1 .. w0 + 3) once and slides them over the KW taps: KW 128-bit
// weight loads and KW + 3 cached global loads feed 16 * KW FMAs (the one-output-per-thread kernel does one
// load per FMA).  Coordinates are the causal frame's; the reflection of an orientation is applied where
// memory is touched.
template <int KWT>
__global__ void __launch_bounds__(256)
conv_wide_kernel(const float *__restrict__ y, const float *__restrict__ weight, float *__restrict__ x,
                 int B, int C, int H, int W, int KH, int Cw, int Cg, int CgP4, int nsplit, int orient)
{
    constexpr int KW = KWT;
    extern __shared__ __align__(16) float wT[];              // [K][Cg][CgP4]
    const int HW = H * W, K = KH * KW;
    const int G = blockIdx.y, tid = threadIdx.x;
    const size_t tap_stride = (size_t)K, row_stride = (size_t)Cw * tap_stride;
    const float *wg = weight + (size_t)G * Cg * row_stride;
    for (int e = tid; e < K * Cg * CgP4; e += blockDim.x) {
        const int co = e % CgP4, ci = (e / CgP4) % Cg, t = e / (CgP4 * Cg);
        const int qh = t / KW, qw = t - qh * KW;
        const int a = (KH - 1 - qh) * KW + (KW - 1 - qw);
        float v = 0.f;
        if (co < Cg) {
            if (t == 0) v = ci < co ? __ldg(wg + co * row_stride + ci * tap_stride + a) : (ci == co ? 1.f : 0.f);
            else        v = __ldg(wg + co * row_stride + ci * tap_stride + a);
        }
        wT[e] = v;
    }
    __syncthreads();

    const bool fw = orient & 1, fh = orient & 2;
    const int nquad = CgP4 >> 2, wq = (W + 3) >> 2;
    const int items = H * wq * nquad;
    for (int u = blockIdx.x; u < B * nsplit; u += gridDim.x) {
        const int b = u / nsplit, sp = u - b * nsplit;
        const float *yb = y + ((size_t)b * C + (size_t)G * Cg) * HW;
        float *xb = x + ((size_t)b * C + (size_t)G * Cg) * HW;
        for (int item = sp * blockDim.x + tid; item < items; item += nsplit * blockDim.x) {
            // consecutive threads -> consecutive pixel runs of one output quad (coalesced y, broadcast weights)
            const int cq = item / (H * wq), r = item - cq * (H * wq);
            const int h = r / wq, w0 = (r - h * wq) * 4;
            float acc[4][4];
#pragma unroll
            for (int px = 0; px < 4; px++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[px][j] = 0.f;
            const int qh_max = h < KH - 1 ? h : KH - 1;
            for (int qh = 0; qh <= qh_max; qh++) {
                const int hn = h - qh, hm = fh ? H - 1 - hn : hn;
                const float *yrow = yb + hm * W;
                const float *wrow = wT + (size_t)(qh * KW) * Cg * CgP4 + cq * 4;
                for (int ci = 0; ci < Cg; ci++) {
                    float yv[KW + 3];                         // y(hn, w0 - KW + 1 + j)
#pragma unroll
                    for (int j = 0; j < KW + 3; j++) {
                        const int w = w0 - (KW - 1) + j;
                        yv[j] = (w >= 0 && w < W) ? __ldg(yrow + (size_t)ci * HW + (fw ? W - 1 - w : w)) : 0.f;
                    }
#pragma unroll
                    for (int qw = 0; qw < KW; qw++) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(wrow + ((size_t)qw * Cg + ci) * CgP4);
#pragma unroll
                        for (int px = 0; px < 4; px++) {
                            const float v = yv[px + (KW - 1) - qw];
                            acc[px][0] = fmaf(w4.x, v, acc[px][0]);
                            acc[px][1] = fmaf(w4.y, v, acc[px][1]);
                            acc[px][2] = fmaf(w4.z, v, acc[px][2]);
                            acc[px][3] = fmaf(w4.w, v, acc[px][3]);
                        }
                    }
                }
            }
            const int co = cq * 4, hmo = fh ? H - 1 - h : h;
#pragma unroll
            for (int px = 0; px < 4; px++) {
                const int w = w0 + px;
                if (w < W) {
                    float *xp = xb + hmo * W + (fw ? W - 1 - w : w);
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (co + j < Cg) xp[(size_t)(co + j) * HW] = acc[px][j];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
conv_kernel(const float *__restrict__ y, const float *__restrict__ weight, float *__restrict__ x,
            int B, int C, int H, int W, int KH, int KW, int Cw, int Cg, int orient)
{
    const int HW = H * W;
    const bool fw = orient & 1, fh = orient & 2;
    const int step_h = fh ? -W : W, step_w = fw ? -1 : 1;
    const size_t total = (size_t)B * C * HW;
    const size_t tap_stride = (size_t)KH * KW, row_stride = (size_t)Cw * tap_stride;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e % HW);
        const int c = (int)((e / HW) % C);
        const size_t b = e / ((size_t)HW * C);
        const int hm = r / W, wm = r - hm * W;
        const int h = fh ? H - 1 - hm : hm, w = fw ? W - 1 - wm : wm;
        const int base = (c / Cg) * Cg, cl = c - base;
        const float *yb = y + (b * C + base) * HW;
        const float *wr = weight + (size_t)c * row_stride;
        float acc = yb[cl * HW + r];
        const int qh_max = h < KH - 1 ? h : KH - 1, qw_max = w < KW - 1 ? w : KW - 1;
        for (int qh = 0; qh <= qh_max; qh++)
            for (int qw = 0; qw <= qw_max; qw++) {
                const int a = (KH - 1 - qh) * KW + (KW - 1 - qw);
                const int rn = r - qh * step_h - qw * step_w;
                const int kc_end = (qh == 0 && qw == 0) ? cl : Cg;   // strictly lower centre tap
                for (int kc = 0; kc < kc_end; kc++)
                    acc = fmaf(__ldg(wr + kc * tap_stride + a), __ldg(yb + kc * HW + rn), acc);
            }
        x[e] = acc;
    }
}

int launch_conv(const Geometry &g, const float *y, const float *weight, float *x, cudaStream_t s)
{
    const size_t total = (size_t)g.B * g.C * g.H * g.W;
    if (total == 0) return 0;
    const int CgP4 = round_up(g.Cg, 4);
    const size_t smem = (size_t)g.K * g.Cg * CgP4 * sizeof(float);
    if (g.Cg >= 3 && g.Cg <= 16 && smem <= (size_t)kMaxSmemBytes) {      // small groups: weights cheap to stage
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(conv_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        int per_sm = (int)((size_t)kMaxSmemBytes / (smem + 1024));
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        int grid_x = (kNumSM * per_sm + g.groups - 1) / g.groups;
        const int items = g.H * g.W * (CgP4 >> 2);
        int nsplit = (grid_x + g.B - 1) / g.B;                 // fill the GPU when the batch is small
        const int max_split = (items + 255) / 256;
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
        if (grid_x > g.B * nsplit) grid_x = g.B * nsplit;
        dim3 grid(grid_x, g.groups);
        conv_tiled_kernel<<<grid, 256, smem, s>>>(y, weight, x, g.B, g.C, g.H, g.W, g.KH, g.KW, g.Cw, g.Cg, CgP4,
                                                  nsplit, g.orient);
        return cuda_status(cudaGetLastError());
    }
    // (staging the weights per CTA only pays when a CTA has enough pixels to spread it over: measured
    //  48.9 vs 24.9 us at (100,48,4,4), 24.8 vs 46.2 us at (256,24,8,8), 335 vs 1119 us at (512,48,16,16))
    bool enough_pixels = (size_t)g.B * g.H * g.W >= (size_t)64 * kNumSM;
    if (env().conv_wide >= 0) enough_pixels = env().conv_wide == 1;      // tests: pin / forbid the wide kernel
    if (g.Cg > 16 && enough_pixels && smem <= (size_t)kMaxSmemBytes &&
        (g.KW == 2 || g.KW == 3 || g.KW == 5 || g.KW == 7)) {
        auto kern = g.KW == 2 ? conv_wide_kernel<2> : g.KW == 3 ? conv_wide_kernel<3>
                  : g.KW == 5 ? conv_wide_kernel<5> : conv_wide_kernel<7>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        int per_sm = (int)((size_t)kMaxSmemBytes / (smem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        int grid_x = (kNumSM * per_sm + g.groups - 1) / g.groups;
        const int items = g.H * ((g.W + 3) / 4) * (CgP4 >> 2);
        int nsplit = (grid_x + g.B - 1) / g.B;                 // fill the GPU when the batch is small
        const int max_split = (items + 255) / 256;
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
        if (grid_x > g.B * nsplit) grid_x = g.B * nsplit;
        dim3 grid(grid_x, g.groups);
        kern<<<grid, 256, smem, s>>>(y, weight, x, g.B, g.C, g.H, g.W, g.KH, g.Cw, g.Cg, CgP4, nsplit, g.orient);
        return cuda_status(cudaGetLastError());
    }
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)kNumSM * 32;
    if (blocks > cap) blocks = cap;
    conv_kernel<<<(unsigned)blocks, 256, 0, s>>>(y, weight, x, g.B, g.C, g.H, g.W, g.KH, g.KW, g.Cw, g.Cg, g.orient);
    return cuda_status(cudaGetLastError());
}

}  // namespace ifk
