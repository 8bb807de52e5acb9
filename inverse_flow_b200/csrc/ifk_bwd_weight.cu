// dW[c][kc][KH-1-qh][KW-1-qw] = - sum_{b,p} dX[b,c,p] * y[b,base+kc,p-q]
//
// Replaces inv_conv_dw (inf/utils/inv_conv_cuda/inv_conv_with_bp_kernel_general.cu:496-735:
// a serial H*W raster loop inside each of <= K*4*C threads, 2*(2K-1)*C/4 launches) by a
// batch-summed correlation of the input gradient with the saved output.
//
// Stage 1  bwd_weight_partial_kernel: CTA = (batch chunk, group, slab of work items); an
//          item is one tap q with a 4x4 tile of (c, kc) and is owned by ONE WARP whose lanes
//          stride over the pixels (conflict-free shared-memory reads), accumulators live in
//          registers across the whole chunk, one shuffle reduction at the end.
// Stage 2  bwd_weight_reduce_kernel: sums the per-chunk partials in chunk order (fixed order,
//          no atomics -> bit-reproducible), negates, masks, scatters into the weight layout.
#include "ifk_internal.cuh"

namespace ifk {

constexpr int kTile = 4;            // (c, kc) register tile edge
constexpr int kItemsPerCta = 16;    // warps per CTA

struct BwdWeightPlan {
    int nt;            // tiles per channel axis
    int items;         // K * nt * nt
    int nz;            // item slabs
    int nchunks;       // batch chunks (== partial buffers)
    int per_chunk;     // images per chunk
    bool staged;       // images resident in shared memory
    int HP, WP, YS;    // halo-padded geometry of the staged y image, its channel stride
    size_t smem_bytes;
};

static BwdWeightPlan make_plan(const Geometry &g)
{
    BwdWeightPlan pl{};
    pl.nt = (g.Cg + kTile - 1) / kTile;
    pl.items = g.K * pl.nt * pl.nt;
    pl.nz = (pl.items + kItemsPerCta - 1) / kItemsPerCta;
    int want = (2 * kNumSM + g.groups * pl.nz - 1) / (g.groups * pl.nz);
    if (want < 1) want = 1;
    if (want > g.B) want = g.B > 0 ? g.B : 1;
    pl.per_chunk = g.B > 0 ? (g.B + want - 1) / want : 1;
    pl.nchunks = g.B > 0 ? (g.B + pl.per_chunk - 1) / pl.per_chunk : 1;
    pl.HP = g.H + g.KH - 1;
    pl.WP = g.W + g.KW - 1;
    pl.YS = pl.HP * pl.WP;
    pl.smem_bytes = (size_t)g.Cg * ((size_t)g.H * g.W + pl.YS) * sizeof(float);
    pl.staged = pl.smem_bytes <= (size_t)kMaxSmemBytes;
    if (!pl.staged) pl.smem_bytes = 0;
    return pl;
}

size_t bwd_weight_workspace_bytes(const Geometry &g)
{
    const BwdWeightPlan pl = make_plan(g);
    return (size_t)pl.nchunks * g.C * g.Cg * g.K * sizeof(float);
}

template <bool STAGED>
__global__ void __launch_bounds__(kItemsPerCta * 32)
bwd_weight_partial_kernel(const float *__restrict__ dx, const float *__restrict__ y,
                          float *__restrict__ partial, int B, int C, int H, int W, int KH, int KW,
                          int Cg, int nt, int items, int per_chunk, int WP, int YS)
{
    extern __shared__ __align__(16) float smem[];
    const int HW = H * W, K = KH * KW;
    const int chunk = blockIdx.x, G = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.z * kItemsPerCta + warp;
    const bool live = item < items;
    const int t = live ? item / (nt * nt) : 0;
    const int rem = live ? item - t * nt * nt : 0;
    const int c0 = (rem / nt) * kTile, k0 = (rem % nt) * kTile;
    const int qh = t / KW, qw = t - qh * KW;
    const int halo = (KH - 1) * WP + (KW - 1);

    float *dxs = smem;                 // [Cg][HW]
    float *ys = smem + Cg * HW;        // [Cg][YS], zero halo on top/left
    if (STAGED) {
        for (int i = threadIdx.x; i < Cg * YS; i += blockDim.x) ys[i] = 0.f;
    }

    float acc[kTile][kTile];
#pragma unroll
    for (int i = 0; i < kTile; i++)
#pragma unroll
        for (int j = 0; j < kTile; j++) acc[i][j] = 0.f;

    // channel indices clamped into the group; out-of-range tile rows are dropped at the end
    int cidx[kTile], kidx[kTile];
#pragma unroll
    for (int i = 0; i < kTile; i++) {
        cidx[i] = c0 + i < Cg ? c0 + i : Cg - 1;
        kidx[i] = k0 + i < Cg ? k0 + i : Cg - 1;
    }

    const int b_begin = chunk * per_chunk;
    const int b_end = b_begin + per_chunk < B ? b_begin + per_chunk : B;
    for (int b = b_begin; b < b_end; b++) {
        const size_t gbase = ((size_t)b * C + (size_t)G * Cg) * HW;
        if (STAGED) {
            __syncthreads();           // previous image fully consumed
            for (int i = threadIdx.x; i < Cg * HW; i += blockDim.x) {
                const int ci = i / HW, r = i - ci * HW;
                const int h = r / W, w = r - h * W;
                dxs[i] = __ldg(dx + gbase + i);
                ys[ci * YS + h * WP + w + halo] = __ldg(y + gbase + i);
            }
            __syncthreads();
            if (live) {
                for (int r = lane; r < HW; r += 32) {
                    const int h = r / W, w = r - h * W;
                    const int yo = h * WP + w + halo - qh * WP - qw;
                    float a[kTile], v[kTile];
#pragma unroll
                    for (int i = 0; i < kTile; i++) {
                        a[i] = dxs[cidx[i] * HW + r];
                        v[i] = ys[kidx[i] * YS + yo];
                    }
#pragma unroll
                    for (int i = 0; i < kTile; i++)
#pragma unroll
                        for (int j = 0; j < kTile; j++) acc[i][j] = fmaf(a[i], v[j], acc[i][j]);
                }
            }
        } else if (live) {
            const float *dxb = dx + gbase, *yb = y + gbase;
            for (int r = lane; r < HW; r += 32) {
                const int h = r / W, w = r - h * W;
                if (h < qh || w < qw) continue;
                const int rn = r - qh * W - qw;
                float a[kTile], v[kTile];
#pragma unroll
                for (int i = 0; i < kTile; i++) {
                    a[i] = __ldg(dxb + (size_t)cidx[i] * HW + r);
                    v[i] = __ldg(yb + (size_t)kidx[i] * HW + rn);
                }
#pragma unroll
                for (int i = 0; i < kTile; i++)
#pragma unroll
                    for (int j = 0; j < kTile; j++) acc[i][j] = fmaf(a[i], v[j], acc[i][j]);
            }
        }
    }

    if (!live) return;
#pragma unroll
    for (int i = 0; i < kTile; i++)
#pragma unroll
        for (int j = 0; j < kTile; j++) {
            float s = acc[i][j];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            acc[i][j] = s;
        }
    if (lane == 0) {
        float *out = partial + ((size_t)chunk * C + (size_t)G * Cg) * Cg * K;   // [c][kc][t]
#pragma unroll
        for (int i = 0; i < kTile; i++)
#pragma unroll
            for (int j = 0; j < kTile; j++)
                if (c0 + i < Cg && k0 + j < Cg)
                    out[((size_t)(c0 + i) * Cg + (k0 + j)) * K + t] = acc[i][j];
    }
}

__global__ void __launch_bounds__(256)
bwd_weight_reduce_kernel(const float *__restrict__ partial, float *__restrict__ dw, int nchunks,
                         int C, int Cg, int Cw, int KH, int KW)
{
    const int K = KH * KW;
    const int total = C * Cw * K;
    const size_t chunk_stride = (size_t)C * Cg * K;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int a = e % K;
        const int kc = (e / K) % Cw;
        const int c = e / (K * Cw);
        const int ah = a / KW, aw = a - ah * KW;
        const int qh = KH - 1 - ah, qw = KW - 1 - aw;
        const int cl = c % Cg;
        float s = 0.f;
        if (kc < Cg && !(qh == 0 && qw == 0 && kc >= cl)) {
            const float *src = partial + ((size_t)c * Cg + kc) * K + (qh * KW + qw);
            for (int n = 0; n < nchunks; n++) s += src[n * chunk_stride];
            s = -s;
        }
        dw[e] = s;
    }
}

int launch_bwd_weight(const Geometry &g, const float *dx, const float *y, float *dw,
                      void *workspace, cudaStream_t s)
{
    const BwdWeightPlan pl = make_plan(g);
    float *partial = (float *)workspace;
    const int total = g.C * g.Cw * g.K;
    if (g.B > 0) {
        dim3 grid(pl.nchunks, g.groups, pl.nz);
        const int threads = kItemsPerCta * 32;
        if (pl.staged) {
            auto kern = bwd_weight_partial_kernel<true>;
            if (pl.smem_bytes > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)pl.smem_bytes);
                if (e != cudaSuccess) return (int)e;
            }
            kern<<<grid, threads, pl.smem_bytes, s>>>(dx, y, partial, g.B, g.C, g.H, g.W, g.KH, g.KW,
                                                      g.Cg, pl.nt, pl.items, pl.per_chunk, pl.WP, pl.YS);
        } else {
            bwd_weight_partial_kernel<false><<<grid, threads, 0, s>>>(
                dx, y, partial, g.B, g.C, g.H, g.W, g.KH, g.KW, g.Cg, pl.nt, pl.items, pl.per_chunk,
                pl.WP, pl.YS);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    int blocks = (total + 255) / 256;
    if (blocks > kNumSM * 8) blocks = kNumSM * 8;
    bwd_weight_reduce_kernel<<<blocks, 256, 0, s>>>(partial, dw, g.B > 0 ? pl.nchunks : 0, g.C, g.Cg,
                                                    g.Cw, g.KH, g.KW);
    return cuda_status(cudaGetLastError());
}

}  // namespace ifk
