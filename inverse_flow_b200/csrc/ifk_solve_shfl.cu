// Register-resident wavefront solve for small images with narrow channel groups ("shuffle" kernel).
//
// The resident kernel (ifk_solve_kernel.cuh) pays, per anti-diagonal, a round trip through shared
// memory: st.shared -> barrier -> ld.shared -> FMA chain -> shuffle reduce (~195 cycles at
// (100,4,14,14) k=2, measured).  When one image's rows fit the lanes of ONE warp and a lane can
// hold the weights of its outputs, nothing in the loop needs shared memory at all:
//   lane = (image row h, channel tile ct); it walks its row, pixel (h, d - h) on diagonal d;
//   the neighbours of that pixel are
//     - pixels of its own row, which the lane (and its sibling tiles) produced on earlier steps,
//     - pixels of the rows above, produced by the lanes of those rows 1..KH-1 steps earlier;
//   each lane keeps a sliding window of those values in registers and receives the newest ones
//   with warp shuffles straight from the producing lanes' registers.
// A step is then: shuffle (the only inter-lane latency) -> the FMAs that involve the fresh values
// (everything older was accumulated while the shuffle was in flight) -> add z -> next step; no
// barrier, no shared-memory dependency.  The image lands in shared memory by one TMA bulk copy; z = T x
// of a pixel is formed on the fly (the lane's rows of T in registers, x loaded one step ahead), so there
// is no pre-pass either; y replaces x in place (fire and forget) and leaves by one TMA bulk store.  A CTA
// is a single warp: no block barrier anywhere.  Lanes left of / right of the image produce zeros, which is exactly
// the causal zero padding their consumers need.
//
// Template: CG channels per group, NCT lanes per row (each CG/NCT output channels), KH x KW taps --
// all compile-time so that weights and windows are registers.  Used when H * NCT <= 32 and (CG, KH,
// KW) is instantiated: the MNIST-sized layers of the reference's models (if_glow_mnist 4x14x14 and
// 8x7x7 with k=2, if_cnn_mnist 1x28x28 with k=3) and the groups=4 grouping of the others.
//
// Replaces the same reference loop as the resident kernel
// (inv_conv_with_bp_kernel_general.cu:72-129; adjoint: .cu:388-483).
#include <stdio.h>
#include <stdlib.h>
#include "ifk_env.cuh"
#include "ifk_solve_kernel.cuh"

namespace ifk {

struct ShflParams {
    const float *in;
    float *out;
    const float *prep;
    int B, C, H, W, KDP, CgP4, XN;
    int flip;           // as SolveParams::flip
    int bulk;
    int early;          // IFK_FLAG_STABLE_PREPARED: weights may be fetched ahead of griddepcontrol.wait
    long long *probe;   // tuning aid: clock64() stamps of CTA (0,0) lane 0 (same slots as the resident kernel)
};

// One "element" of the input-channel axis: a packed pair of channels (two FMAs per instruction, SASS
// FFMA2) when the group width is even, a single float otherwise.  A lone warp issues in order and a
// three-register FFMA occupies the FMA pipe for two cycles, so halving the FMA count shortens the step.
template <int P> struct ShflElem;
template <> struct ShflElem<1> {
    typedef float type;
    __device__ __forceinline__ static float zero() { return 0.f; }
    __device__ __forceinline__ static float make(const float *v) { return v[0]; }
    __device__ __forceinline__ static float fma(float a, float b, float c) { return fmaf(a, b, c); }
    __device__ __forceinline__ static float sum(float v) { return v; }
};
template <> struct ShflElem<2> {
    typedef f32x2_t type;
    __device__ __forceinline__ static f32x2_t zero() { return 0ull; }
    __device__ __forceinline__ static f32x2_t make(const float *v) { return pack_f32x2(v[0], v[1]); }
    __device__ __forceinline__ static f32x2_t fma(f32x2_t a, f32x2_t b, f32x2_t c) { return fma_f32x2(a, b, c); }
    __device__ __forceinline__ static float sum(f32x2_t v) { return sum_f32x2(v); }
};

template <int CG, int NCT, int KH, int KW>
__global__ void __launch_bounds__(32)
solve_shfl_kernel(const ShflParams p)
{
    constexpr int CC = CG / NCT;            // output channels per lane
    constexpr int K = KH * KW;
    constexpr int KW1 = KW > 1 ? KW - 1 : 1;
    // Packed pairs (P = 2) were measured SLOWER here than scalar FMAs (174 vs 152 cycles per diagonal at
    // (100,4,14,14) k=2): the pack moves and the longer FFMA2 latency sit on the one warp's dependent chain.
    constexpr int P = 1;                                // input channels per element
    constexpr int CE = CG / P;                          // elements per pixel
    typedef ShflElem<P> E;
    typedef typename E::type elem_t;
    static_assert(CG % NCT == 0, "channel tiles must divide the group");
    IFK_PROBE(0);
    extern __shared__ __align__(128) float smem[];
    const int H = p.H, W = p.W, HW = p.H * p.W;
    const int lane = threadIdx.x;

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);         // 16 bytes reserved
    float *xbuf = smem + 4;                                     // [CG][HW]: x lands here, y replaces it in place

    const int G = blockIdx.y;
    const float *wg = p.prep + (size_t)G * CG * p.KDP;
    const uint32_t img_bytes = (uint32_t)(CG * HW) * 4u;
    const size_t img_stride = (size_t)p.C * HW;
    const float *in0 = p.in + (size_t)G * CG * HW;
    float *out0 = p.out + (size_t)G * CG * HW;

    // programmatic dependent launch: as in the resident kernel, this lane's weights and rows of T are
    // fetched ahead of the dependency wait only under IFK_FLAG_STABLE_PREPARED (ifk.h)
    if (!p.early) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }

    const int row = lane / NCT, ct = lane - row * NCT;
    const bool valid = row < H;
    elem_t wreg[CC][K > 1 ? K - 1 : 1][CE];
    elem_t treg[CC][CE];                    // rows of T = (I + A0)^-1: z = T x is formed on the fly
#pragma unroll
    for (int i = 0; i < CC; i++) {
        const float *wrow = wg + (size_t)(ct * CC + i) * p.KDP;
#pragma unroll
        for (int t = 1; t < K; t++)
#pragma unroll
            for (int e = 0; e < CE; e++) {
                float tmp[P];
#pragma unroll
                for (int q = 0; q < P; q++) tmp[q] = __ldg(wrow + t * CG + e * P + q);
                wreg[i][t - 1][e] = E::make(tmp);
            }
#pragma unroll
        for (int e = 0; e < CE; e++) {
            float tmp[P];
#pragma unroll
            for (int q = 0; q < P; q++) tmp[q] = __ldg(wrow + e * P + q);
            treg[i][e] = E::make(tmp);
        }
    }
    if (p.early) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    int b = blockIdx.x;
    if (p.bulk && lane == 0) {
        mbar_init(bar, 1);
        if (b < p.B) {
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b * img_stride, img_bytes, bar);
        }
    }
    __syncwarp();
    IFK_PROBE(1);       // weights and T in registers, previous kernel finished (griddepcontrol.wait), load issued
    IFK_PROBE(2);
    IFK_PROBE(3);

    // memory index of solver pixel (h, w) = idx0 + sh*h*W + sw*w (reflected axes walk backwards)
    const int sw = (p.flip & 1) ? -1 : 1, sh = (p.flip & 2) ? -1 : 1;
    const int idx0 = ((p.flip & 2) ? (H - 1) * W : 0) + ((p.flip & 1) ? W - 1 : 0);
    const int ndiag = H + W - 1;
    // pixel (row, d - row) of channel 0: x is read from every channel, y written to channels ct*CC + i
    const uint32_t x_lane0 = smem_u32(xbuf) + (uint32_t)(idx0 + (sh * W - sw) * (valid ? row : 0)) * 4u;
    const uint32_t x_step = (uint32_t)(sw * 4);
    const uint32_t ch_stride = (uint32_t)HW * 4u;
    const uint32_t y_off = (uint32_t)(ct * CC) * ch_stride;
    const uint32_t scratch = smem_u32(smem + 3);                // unused word of the 16 reserved bytes

    uint32_t parity = 0;
    for (; b < p.B; b += gridDim.x) {
        const int b_next = b + gridDim.x;
        if (p.bulk) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            const float *src = in0 + (size_t)b * img_stride;
            for (int i = lane; i < CG * HW; i += 32) xbuf[i] = __ldg(src + i);
            __syncwarp();
        }
        IFK_PROBE(4);   // image landed
        IFK_PROBE(5);

        {
            // win[qh][j] = y(h - qh, w - 1 - j): what this step needs at column offset j + 1
            elem_t win[KH][KW1][CE];
            float hist[KH > 1 ? KH - 1 : 1][CC];    // this lane's outputs 2, 3, .. steps ago
#pragma unroll
            for (int qh = 0; qh < KH; qh++)
#pragma unroll
                for (int j = 0; j < KW1; j++)
#pragma unroll
                    for (int c = 0; c < CE; c++) win[qh][j][c] = E::zero();
#pragma unroll
            for (int s = 0; s < (KH > 1 ? KH - 1 : 1); s++)
#pragma unroll
                for (int i = 0; i < CC; i++) hist[s][i] = 0.f;
            float ylast[CC];                        // this lane's output of the previous step
#pragma unroll
            for (int i = 0; i < CC; i++) ylast[i] = 0.f;

            uint32_t xa = x_lane0;                  // pixel (row, d - row), channel 0
            int col = -row;
            float xv[CG];                           // x of this step's pixel, loaded one step ahead
#pragma unroll
            for (int c = 0; c < CG; c++) xv[c] = (valid && col == 0) ? lds_f32(xa + ch_stride * c) : 0.f;

            for (int d = 0; d < ndiag; d++, col++, xa += x_step) {
                const bool active = valid && (unsigned)col < (unsigned)W;
                // 0. x of the NEXT step's pixel: issued first, so that the loads have the whole step to land
                const bool next_active = valid && (unsigned)(col + 1) < (unsigned)W;
                float xnext[CG];
#pragma unroll
                for (int c = 0; c < CG; c++) xnext[c] = next_active ? lds_f32(xa + x_step + ch_stride * c) : 0.f;
                // 1. the newest neighbours, straight from the producing lanes' registers
                float fresh[KH][CG];                // fresh[qh] = y(h - qh, w) for qh >= 1; fresh[0] = y(h, w - 1)
#pragma unroll
                for (int c2 = 0; c2 < NCT; c2++)
#pragma unroll
                    for (int i = 0; i < CC; i++) {
                        if (NCT == 1) fresh[0][i] = ylast[i];
                        else fresh[0][c2 * CC + i] = __shfl_sync(0xffffffffu, ylast[i], row * NCT + c2);
                    }
#pragma unroll
                for (int qh = 1; qh < KH; qh++)
#pragma unroll
                    for (int c2 = 0; c2 < NCT; c2++)
#pragma unroll
                        for (int i = 0; i < CC; i++) {
                            // the output of lane (row - qh) qh steps ago: ylast = 1 step, hist[0] = 2 steps, ...
                            const float src_v = qh == 1 ? ylast[i] : hist[qh - 2][i];
                            const int src = row >= qh ? (row - qh) * NCT + c2 : lane;
                            const float got = __shfl_sync(0xffffffffu, src_v, src);
                            fresh[qh][c2 * CC + i] = row >= qh ? got : 0.f;
                        }

                // 2. while the shuffles are in flight: z = T x of this pixel and the older columns
                elem_t acc[CC], late[CC], xe[CE];
#pragma unroll
                for (int e = 0; e < CE; e++) xe[e] = E::make(xv + e * P);
#pragma unroll
                for (int i = 0; i < CC; i++) {
                    acc[i] = E::zero();
                    late[i] = E::zero();
                }
#pragma unroll
                for (int e = 0; e < CE; e++)
#pragma unroll
                    for (int i = 0; i < CC; i++) late[i] = E::fma(treg[i][e], xe[e], late[i]);
#pragma unroll
                for (int qh = 0; qh < KH; qh++)
#pragma unroll
                    for (int qw = (qh == 0 ? 2 : 1); qw < KW; qw++)
#pragma unroll
                        for (int e = 0; e < CE; e++)
#pragma unroll
                            for (int i = 0; i < CC; i++)
                                acc[i] = E::fma(wreg[i][qh * KW + qw - 1][e], win[qh][qw - 1][e], acc[i]);
                //    then the fresh values: tap (0, 1) on one chain, taps (qh, 0) on the other
                elem_t fe[KH][CE];
#pragma unroll
                for (int qh = 0; qh < KH; qh++)
#pragma unroll
                    for (int e = 0; e < CE; e++) fe[qh][e] = E::make(fresh[qh] + e * P);
#pragma unroll
                for (int e = 0; e < CE; e++)
#pragma unroll
                    for (int i = 0; i < CC; i++) {
                        if (KW > 1) late[i] = E::fma(wreg[i][0][e], fe[0][e], late[i]);
#pragma unroll
                        for (int qh = 1; qh < KH; qh++)
                            acc[i] = E::fma(wreg[i][qh * KW - 1][e], fe[qh][e], acc[i]);
                    }
                float y[CC];
#pragma unroll
                for (int i = 0; i < CC; i++) {
                    y[i] = active ? E::sum(acc[i]) + E::sum(late[i]) : 0.f;
                    // branch-free: lanes off the image write a scratch word (a divergent branch per
                    // output costs more than the whole FMA chain of a step).  y replaces x in place:
                    // every lane of this row read x of this pixel one step ago.  (Storing y straight to
                    // HBM from here instead was measured slower: 4.15 vs 3.48 us per kernel.)
                    sts_f32(active ? xa + y_off + ch_stride * i : scratch, y[i]);
                }
                // 3. slide the windows
#pragma unroll
                for (int c = 0; c < CG; c++) xv[c] = xnext[c];
                // windows for the next step (column w + 1): win'[qh][j] = y(h - qh, w - j).  Row 0 lags:
                // y(h, w) of this step reaches the sibling lanes by next step's shuffle (fresh[0]), so
                // win[0][0] is never read and win'[0][1] = y(h, w - 1) = this step's fresh[0].
#pragma unroll
                for (int qh = 1; qh < KH; qh++) {
#pragma unroll
                    for (int j = KW1 - 1; j >= 1; j--)
#pragma unroll
                        for (int c = 0; c < CE; c++) win[qh][j][c] = win[qh][j - 1][c];
#pragma unroll
                    for (int c = 0; c < CE; c++) win[qh][0][c] = fe[qh][c];
                }
#pragma unroll
                for (int j = KW1 - 1; j >= 2; j--)
#pragma unroll
                    for (int c = 0; c < CE; c++) win[0][j][c] = win[0][j - 1][c];
                if (KW1 > 1) {
#pragma unroll
                    for (int c = 0; c < CE; c++) win[0][1][c] = fe[0][c];
                }
#pragma unroll
                for (int s = (KH > 1 ? KH - 2 : 0); s >= 1; s--)
#pragma unroll
                    for (int i = 0; i < CC; i++) hist[s][i] = hist[s - 1][i];
                if (KH > 2) {
#pragma unroll
                    for (int i = 0; i < CC; i++) hist[0][i] = ylast[i];
                }
#pragma unroll
                for (int i = 0; i < CC; i++) ylast[i] = y[i];
            }
        }

        IFK_PROBE(6);   // diagonal loop done
        float *dst = out0 + (size_t)b * img_stride;
        if (p.bulk) {
            fence_async_proxy();            // generic-proxy writes of xbuf -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
                bulk_store(dst, xbuf, img_bytes);
                if (b_next < p.B) {         // the buffer is reused: only after the store has read it
                    bulk_store_wait_read();
                    mbar_expect_tx(bar, img_bytes);
                    bulk_load(xbuf, in0 + (size_t)b_next * img_stride, img_bytes, bar);
                }
            }
        } else {
            __syncwarp();
            for (int i = lane; i < CG * HW; i += 32) dst[i] = xbuf[i];
            __syncwarp();
        }
    }
    IFK_PROBE(7);
    if (p.bulk && lane == 0) bulk_store_wait_read();   // smem must outlive the last store's read
    IFK_PROBE(8);
}

// ---- host side -----------------------------------------------------------------------------
// X(CG, NCT, KH, KW)
#define IFK_SHFL_VARIANTS                                                                             \
    X(1, 1, 2, 2) X(1, 1, 3, 3) X(2, 1, 2, 2) X(2, 2, 2, 2) X(2, 1, 3, 3) X(3, 1, 3, 3) X(3, 1, 2, 2)  \
    X(4, 1, 2, 2) X(4, 2, 2, 2) X(4, 1, 3, 3) X(4, 2, 3, 3) X(8, 2, 2, 2) X(8, 4, 2, 2)
// (CG = 6 at k = 3 was measured slower than the resident kernel -- 10.0 vs 6.5 us at (100,24,8,8) groups=4:
//  288 weights and a 48-value window per row no longer fit a few lanes' registers comfortably)

struct ShflConfig {
    bool ok;
    int nct, grid_x;
    size_t smem_bytes;
};

static ShflConfig choose_shfl(const Geometry &g)
{
    ShflConfig c{};
    c.ok = false;
    const EnvKnobs &k = env();
    if (k.shfl_off || k.pins_other_solver()) return c;          // tests / tuning runs that pin another kernel
    const int want_nct = k.shfl_nct;                            // tuning only
    const int XN = round_up(g.Cg * g.H * g.W, 4);
    const size_t smem = 16 + (size_t)XN * sizeof(float);
    if (smem > (size_t)kMaxSmemBytes) return c;
    int best_nct = 0;
#define X(CG, NCT, KHc, KWc)                                                                          \
    if (g.Cg == CG && g.KH == KHc && g.KW == KWc && g.H * NCT <= 32 && (!want_nct || want_nct == NCT)) \
        if (NCT > best_nct) best_nct = NCT;
    IFK_SHFL_VARIANTS
#undef X
    if (!best_nct) return c;
    c.ok = true;
    c.nct = best_nct;
    c.smem_bytes = smem;
    int per_sm = (int)((size_t)(kMaxSmemBytes + 1024) / (smem + 1024));
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    int grid_x = (kNumSM * per_sm + g.groups - 1) / g.groups;
    if (grid_x > g.B) grid_x = g.B;
    if (grid_x < 1) grid_x = 1;
    c.grid_x = grid_x;
    return c;
}

bool shfl_solve_available(const Geometry &g) { return choose_shfl(g).ok; }

int describe_shfl_solve(const Geometry &g, char *buf, size_t buflen)
{
    const ShflConfig c = choose_shfl(g);
    snprintf(buf, buflen, "shfl<cg=%d,nct=%d,k=%dx%d> lanes=%d threads=32 smem=%zuB grid=%dx%d", g.Cg, c.nct,
             g.KH, g.KW, g.H * c.nct, c.smem_bytes, c.grid_x, g.groups);
    return 0;
}

int launch_solve_shfl(const Geometry &g, const float *in, const float *prep_dir, float *out, bool reverse,
                      long long *probe, cudaStream_t s)
{
    const ShflConfig c = choose_shfl(g);
    if (!c.ok) return IFK_ERR_UNSUPPORTED;
    ShflParams p{};
    p.in = in; p.out = out; p.prep = prep_dir;
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W; p.KDP = g.KDP;
    p.CgP4 = round_up(g.Cg, 4);
    p.XN = round_up(g.Cg * g.H * g.W, 4);
    p.flip = reverse ? (g.orient ^ 3) : g.orient;
    const size_t img_bytes = (size_t)g.Cg * g.H * g.W * sizeof(float);
    p.bulk = (img_bytes % 16 == 0) && (((uintptr_t)in | (uintptr_t)out) % 16 == 0) ? 1 : 0;
    if (env().nobulk) p.bulk = 0;
    p.early = (g.flags & IFK_FLAG_STABLE_PREPARED) ? 1 : 0;
    p.probe = probe;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c.grid_x, g.groups);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = c.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL, see the kernel prologue
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = solve_use_pdl() ? 1 : 0;
#define X(CG, NCT, KHc, KWc)                                                                          \
    if (g.Cg == CG && c.nct == NCT && g.KH == KHc && g.KW == KWc) {                                   \
        auto kern = solve_shfl_kernel<CG, NCT, KHc, KWc>;                                             \
        if (c.smem_bytes > 48 * 1024) {                                                               \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 (int)c.smem_bytes);                                  \
            if (e != cudaSuccess) return (int)e;                                                      \
        }                                                                                             \
        return cuda_status(cudaLaunchKernelEx(&cfg, kern, p));                                        \
    }
    IFK_SHFL_VARIANTS
#undef X
    return IFK_ERR_UNSUPPORTED;
}

}  // namespace ifk
