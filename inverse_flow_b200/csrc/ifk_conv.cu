// x = L y : the masked (autoregressive) convolution, sampling direction z -> x.
// Replaces inv_conv_fwd_cuda_inverse (inv_conv_with_bp_kernel_general.cu:141-264), which
// walks the anti-diagonals with a launch + device sync each although nothing depends on
// anything here.  One launch, one thread per output element, coalesced along W; the centre
// tap is masked to its strictly-lower triangle and the diagonal is the implicit 1.
#include "ifk_internal.cuh"

namespace ifk {

__global__ void __launch_bounds__(256)
conv_kernel(const float *__restrict__ y, const float *__restrict__ weight, float *__restrict__ x,
            int B, int C, int H, int W, int KH, int KW, int Cw, int Cg)
{
    const int HW = H * W;
    const size_t total = (size_t)B * C * HW;
    const size_t tap_stride = (size_t)KH * KW, row_stride = (size_t)Cw * tap_stride;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e % HW);
        const int c = (int)((e / HW) % C);
        const size_t b = e / ((size_t)HW * C);
        const int h = r / W, w = r - h * W;
        const int base = (c / Cg) * Cg, cl = c - base;
        const float *yb = y + (b * C + base) * HW;
        const float *wr = weight + (size_t)c * row_stride;
        float acc = yb[cl * HW + r];
        const int qh_max = h < KH - 1 ? h : KH - 1, qw_max = w < KW - 1 ? w : KW - 1;
        for (int qh = 0; qh <= qh_max; qh++)
            for (int qw = 0; qw <= qw_max; qw++) {
                const int a = (KH - 1 - qh) * KW + (KW - 1 - qw);
                const int rn = r - qh * W - qw;
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
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)kNumSM * 32;
    if (blocks > cap) blocks = cap;
    conv_kernel<<<(unsigned)blocks, 256, 0, s>>>(y, weight, x, g.B, g.C, g.H, g.W, g.KH, g.KW, g.Cw, g.Cg);
    return cuda_status(cudaGetLastError());
}

}  // namespace ifk
