// dW[c][kc][KH-1-qh][KW-1-qw] = - sum_{b,p} dX[b,c,p] * y[b,base+kc,p-q]
//
// Replaces inv_conv_dw (inf/utils/inv_conv_cuda/inv_conv_with_bp_kernel_general.cu:496-735:
// a serial H*W raster loop inside each of <= K*4*C threads, 2*(2K-1)*C/4 launches) by a
// batch-summed correlation of the input gradient with the saved output.
//
// Stage 1  bwd_weight_partial_kernel: CTA = (batch chunk, group, slab of work items).  A work
//          item is one tap q with a TC x TK tile of (c, kc) (4 x 12 for Cg >= 12); a warp owns
//          one item and keeps its TC*TK accumulators in registers across the whole chunk; its
//          lanes stride over the pixels, so every shared-memory read is a conflict-free row of
//          32 distinct words and one pixel costs TC+TK loads for TC*TK FMAs.  The chunk's images (dX and y, both contiguous NCHW group slices)
//          arrive by TMA bulk copies, double buffered: image n+1 lands while image n is
//          consumed.  One recursive-halving shuffle reduction at the very end.
// Stage 2  bwd_weight_reduce_kernel: sums the per-chunk partials in chunk order (fixed order,
//          no atomics -> bit-reproducible), negates, masks, scatters into the weight layout.
#include "ifk_env.cuh"
#include "ifk_solve_kernel.cuh"   // TMA / mbarrier primitives

namespace ifk {

constexpr int kWarps = 16;       // warps per CTA; one work item (tap x TC x TK tile) per warp

struct BwdWeightPlan {
    int tc, tk;        // register tile: TC dX channels x TK y channels
    int ntc, ntk;      // tiles per channel axis
    int items;         // K * ntc * ntk
    int nz;            // item slabs (grid.z)
    int nchunks;       // batch chunks (== partial buffers)
    int per_chunk;     // images per chunk
    int nbuf;          // 2: double buffered, 1: single, 0: images read from global memory
    int nstage;        // images staged per buffer (small images travel in batches)
    int CgP;           // channels padded to the tile sizes (padding rows stay zero in smem)
    int XN;            // floats per staged image (16-byte multiple)
    size_t smem_bytes;
};

static BwdWeightPlan make_plan(const Geometry &g)
{
    BwdWeightPlan pl{};
    pl.tc = g.Cg >= 4 ? 4 : (g.Cg >= 2 ? 2 : 1);
    pl.tk = g.Cg >= 12 ? 12 : (g.Cg >= 8 ? 8 : (g.Cg >= 4 ? 4 : pl.tc));
    pl.ntc = (g.Cg + pl.tc - 1) / pl.tc;
    pl.ntk = (g.Cg + pl.tk - 1) / pl.tk;
    pl.items = g.K * pl.ntc * pl.ntk;
    pl.nz = (pl.items + kWarps - 1) / kWarps;
    int want = (kNumSM + g.groups * pl.nz - 1) / (g.groups * pl.nz);
    if (want < 1) want = 1;
    if (want > g.B) want = g.B > 0 ? g.B : 1;
    pl.per_chunk = g.B > 0 ? (g.B + want - 1) / want : 1;
    pl.nchunks = g.B > 0 ? (g.B + pl.per_chunk - 1) / pl.per_chunk : 1;
    const int cp1 = pl.ntc * pl.tc, cp2 = pl.ntk * pl.tk;
    pl.CgP = cp1 > cp2 ? cp1 : cp2;
    pl.XN = round_up(pl.CgP * g.H * g.W, 4);
    const size_t one = (size_t)2 * pl.XN * sizeof(float);          // dX + y of one image
    pl.nbuf = 2 * one + 64 <= (size_t)kMaxSmemBytes ? 2 : (one + 64 <= (size_t)kMaxSmemBytes ? 1 : 0);
    if (pl.per_chunk == 1 && pl.nbuf == 2) pl.nbuf = 1;
    // small images: several per buffer, so that one TMA round trip feeds many pixel steps
    pl.nstage = 1;
    if (pl.nbuf) {
        const size_t budget = 96 * 1024;                           // per buffer
        int n = (int)(budget / one);
        if (n < 1) n = 1;
        if (n > 16) n = 16;
        const int need = (pl.per_chunk + pl.nbuf - 1) / pl.nbuf;   // no point staging more than this
        if (n > need) n = need > 0 ? need : 1;
        pl.nstage = n;
    }
    pl.smem_bytes = pl.nbuf ? 64 + (size_t)pl.nbuf * pl.nstage * one : 0;
    return pl;
}

// batch chunks (== partial buffers) of the stage-1 kernel that serves this geometry
static int stage1_chunks(const Geometry &g);

size_t bwd_weight_workspace_bytes(const Geometry &g)
{
    return (size_t)stage1_chunks(g) * g.C * g.Cg * g.K * sizeof(float);
}

struct BwdWeightParams {
    const float *dx, *y;
    float *partial;
    int B, C, H, W, KH, KW, Cg, ntk, items, per_chunk, nbuf, nstage, XN, bulk, orient;
    unsigned w_magic;   // ceil(2^32 / W): r / W == umulhi(r, w_magic) while r * W < 2^32 (staged images are far
                        // smaller); 0 when W == 1
};

// sum N per-lane values over the 32 lanes by recursive halving: N/2 + N/4 + ... shuffles
// (against 5N for butterflies); lane l ends with the totals of entries [off, off+size),
// (off, size) = rs_owner(N, 32, l), in a[0 .. size)
template <int N, int M>
struct Halve {
    __device__ __forceinline__ static void run(float *a, int lane)
    {
        constexpr int HALF = (N + 1) / 2;
        const bool hi = (lane & (M / 2)) != 0;
#pragma unroll
        for (int i = 0; i < HALF; i++) {
            const float lo_v = a[i];
            const float hi_v = (i + HALF < N) ? a[i + HALF] : 0.f;
            a[i] = (hi ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, hi ? lo_v : hi_v, M / 2);
        }
        Halve<HALF, M / 2>::run(a, lane);
    }
};
template <int N>
struct Halve<N, 1> {
    __device__ __forceinline__ static void run(float *, int) {}
};
__host__ __device__ constexpr int halve_final(int n, int m) { return m > 1 ? halve_final((n + 1) / 2, m / 2) : n; }

template <int TC, int TK>
__global__ void __launch_bounds__(kWarps * 32)
bwd_weight_partial_kernel(const BwdWeightParams p)
{
    extern __shared__ __align__(128) float smem[];
    const int W = p.W, HW = p.H * p.W, K = p.KH * p.KW, Cg = p.Cg;
    const int chunk = blockIdx.x, G = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);          // [2] one per buffer
    float *buf0 = smem + 16;                                      // buffer b: dX at b*2*XN, y behind it

    // this warp's work item: one tap, TC dX channels x TK y channels
    const int item = blockIdx.z * kWarps + warp;
    const bool live = item < p.items;
    const int t = live ? item % K : 0, tile = live ? item / K : 0;
    const int c0 = (tile / p.ntk) * TC, k0 = (tile % p.ntk) * TK;
    const int qh = t / p.KW, qw = t - qh * p.KW;
    // orientation (ifk.h): on a reflected axis the neighbour p - q lies at a larger memory index
    const bool fw = p.orient & 1, fh = p.orient & 2;
    const int aoff = c0 * HW, voff = k0 * HW - ((fh ? -qh : qh) * W + (fw ? -qw : qw));
    const int h_lo = fh ? 0 : qh, h_hi = fh ? p.H - 1 - qh : p.H - 1;     // rows / columns whose neighbour exists
    const int w_lo = fw ? 0 : qw, w_hi = fw ? W - 1 - qw : W - 1;
    // a tap that reaches outside the image from every pixel (kernel larger than the image) sums nothing;
    // otherwise `neutral_r` is a pixel whose neighbour exists: the load address of lanes that contribute zero
    const bool empty = h_lo > h_hi || w_lo > w_hi;
    const int neutral_r = empty ? 0 : h_lo * W + w_lo;
    float acc[TC * TK];
#pragma unroll
    for (int i = 0; i < TC * TK; i++) acc[i] = 0.f;

    const int b_begin = chunk * p.per_chunk;
    const int b_end = b_begin + p.per_chunk < p.B ? b_begin + p.per_chunk : p.B;
    const size_t img_stride = (size_t)p.C * HW;
    const float *dx0 = p.dx + (size_t)G * Cg * HW, *y0 = p.y + (size_t)G * Cg * HW;
    const uint32_t img_bytes = (uint32_t)(Cg * HW) * 4u;
    const bool staged = p.nbuf > 0, bulk = staged && p.bulk;
    const int XN = p.XN, nstage = p.nstage;
    const size_t buf_floats = (size_t)2 * nstage * XN;            // one buffer: nstage x (dX, y)

    // stage `n` images starting at image b into buffer `which` (thread 0; TMA bulk copies)
    auto issue = [&](int which, int b, int n) {
        float *bufw = buf0 + (size_t)which * buf_floats;
        mbar_expect_tx(&bars[which], 2u * img_bytes * (uint32_t)n);
        for (int i = 0; i < n; i++) {
            bulk_load(bufw + (size_t)(2 * i) * XN, dx0 + (size_t)(b + i) * img_stride, img_bytes, &bars[which]);
            bulk_load(bufw + (size_t)(2 * i + 1) * XN, y0 + (size_t)(b + i) * img_stride, img_bytes, &bars[which]);
        }
    };

    if (bulk && tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        const int n0 = b_end - b_begin < nstage ? b_end - b_begin : nstage;
        issue(0, b_begin, n0);
    }
    if (staged && XN > Cg * HW) {   // channels padded up to the tile are zero for good (bulk copies never touch them)
        for (int bsel = 0; bsel < 2 * p.nbuf * nstage; bsel++)
            for (int i = Cg * HW + tid; i < XN; i += blockDim.x) buf0[(size_t)bsel * XN + i] = 0.f;
    }
    __syncthreads();
    uint32_t parity0 = 0u, parity1 = 0u;

    int stage_idx = 0;
    for (int b = b_begin; b < b_end; b += nstage, stage_idx++) {
        const int n_img = b_end - b < nstage ? b_end - b : nstage;
        const int cur = p.nbuf == 2 ? stage_idx & 1 : 0;
        const float *stage_base;
        if (staged) {
            float *bufc = buf0 + (size_t)cur * buf_floats;
            if (bulk) {
                const int b_nx = b + nstage;
                if (p.nbuf == 2 && b_nx < b_end && tid == 0)             // prefetch the next stage
                    issue(cur ^ 1, b_nx, b_end - b_nx < nstage ? b_end - b_nx : nstage);
                mbar_wait(&bars[cur], cur ? parity1 : parity0);
                if (cur) parity1 ^= 1u; else parity0 ^= 1u;
            } else {
                __syncthreads();
                for (int im = 0; im < n_img; im++) {
                    const float *sd = dx0 + (size_t)(b + im) * img_stride, *sy = y0 + (size_t)(b + im) * img_stride;
                    for (int i = tid; i < Cg * HW; i += blockDim.x) {
                        bufc[(size_t)(2 * im) * XN + i] = __ldg(sd + i);
                        bufc[(size_t)(2 * im + 1) * XN + i] = __ldg(sy + i);
                    }
                }
                __syncthreads();
            }
            stage_base = bufc;
        } else {
            stage_base = nullptr;
        }

        if (live && staged && !empty) {
            // lanes stride over the (image, pixel) pairs of the stage: every shared-memory read is 32
            // consecutive words, and tiny images still fill the warp.  The walk advances (image, pixel)
            // by 32 pairs with one conditional subtraction; row / column come from one multiply-high;
            // pixels whose neighbour lies outside the image multiply by zero instead of branching
            // (a divergent branch around the tile costs more than its FMAs at these sizes).
            int im = lane / HW, r = lane - im * HW;          // once per stage
            const int d_im = 32 / HW, d_r = 32 - d_im * HW;
            while (im < n_img) {
                const int h = p.w_magic ? (int)__umulhi((unsigned)r, p.w_magic) : r, w = r - h * W;   // W == 1: h = r
                const bool ok = h >= h_lo && h <= h_hi && w >= w_lo && w <= w_hi;
                const float *dxs = stage_base + (size_t)(2 * im) * XN, *ys = dxs + XN;
                const int rv = ok ? r : neutral_r;           // a pixel whose neighbour exists (address safety)
                float a[TC], v[TK];
#pragma unroll
                for (int i = 0; i < TC; i++) a[i] = ok ? dxs[aoff + i * HW + r] : 0.f;
#pragma unroll
                for (int k = 0; k < TK; k++) v[k] = ys[voff + k * HW + rv];
#pragma unroll
                for (int i = 0; i < TC; i++)
#pragma unroll
                    for (int k = 0; k < TK; k++) acc[i * TK + k] = fmaf(a[i], v[k], acc[i * TK + k]);
                r += d_r;
                im += d_im;
                if (r >= HW) { r -= HW; im++; }
            }
        } else if (live && !staged) {
            int im = 0, h = 0, w = lane;
            while (w >= W) { w -= W; h++; }
            while (h >= p.H) { h -= p.H; im++; }
            while (im < n_img) {
                if (h >= h_lo && h <= h_hi && w >= w_lo && w <= w_hi) {
                    const int r = h * W + w;
                    float a[TC], v[TK];
                    const float *dxs = dx0 + (size_t)(b + im) * img_stride, *ys = y0 + (size_t)(b + im) * img_stride;
#pragma unroll
                    for (int i = 0; i < TC; i++)
                        a[i] = c0 + i < Cg ? __ldg(dxs + aoff + i * HW + r) : 0.f;
#pragma unroll
                    for (int k = 0; k < TK; k++)
                        v[k] = k0 + k < Cg ? __ldg(ys + voff + k * HW + r) : 0.f;
#pragma unroll
                    for (int i = 0; i < TC; i++)
#pragma unroll
                        for (int k = 0; k < TK; k++) acc[i * TK + k] = fmaf(a[i], v[k], acc[i * TK + k]);
                }
                w += 32;
                while (w >= W) { w -= W; h++; }
                while (h >= p.H) { h -= p.H; im++; }
            }
        }
        if (bulk && p.nbuf == 2) __syncthreads();      // everyone is done with `cur` before it is refilled
        if (bulk && p.nbuf == 1 && b + nstage < b_end) {
            __syncthreads();
            if (tid == 0) {
                const int b_nx = b + nstage;
                issue(0, b_nx, b_end - b_nx < nstage ? b_end - b_nx : nstage);
            }
        }
    }

    if (!live) return;                                 // warp-uniform
    Halve<TC * TK, 32>::run(acc, lane);
    int off, size;
    rs_owner(TC * TK, 32, lane, &off, &size);
    float *out = p.partial + ((size_t)chunk * p.C + (size_t)G * Cg) * Cg * K;   // [c][kc][t]
    constexpr int kFinal = halve_final(TC * TK, 32);
#pragma unroll
    for (int i = 0; i < kFinal; i++) {
        const int e = off + i;
        const int ic = e / TK, ik = e - ic * TK;
        if (i < size && c0 + ic < Cg && k0 + ik < Cg)
            out[((size_t)(c0 + ic) * Cg + (k0 + ik)) * K + t] = acc[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 1 for the reference models' layers ("quad" kernel): tiles that span ALL taps, 4 pixels per lane.
//
// The kernel above re-reads dX and y from shared memory once per tap (one work item per tap) and loads
// TC + TK scalars per TC*TK FMAs.  Here a warp owns a tile of TC dX channels x TKC y channels x all K taps
// (72 accumulators for 4 x 2 x 9) and a lane a QUAD of four horizontally adjacent pixels: dX arrives as one
// 128-bit load per channel, the y neighbourhood of the quad as two 128-bit loads per (channel, tap row) --
// an 8-wide window from which every horizontal tap is a register offset -- so one quad costs
// TC + 2*TKC*KH vector loads for 4*TC*TKC*K FMAs (16 : 288).  Lanes stride over the (image, quad) pairs of
// the CTA's batch chunk, which is staged once by TMA bulk copies (no re-staging per tap); one recursive-
// halving reduction per tile at the end.  Grid = (batch chunks, groups, tile groups), sized to about a
// third of the SMs: with 256 threads of ~170 registers a CTA cannot share an SM with a wavefront-solve CTA,
// so in the backward pass these kernels live on the SMs the 100-image solves leave idle.
// ------------------------------------------------------------------------------------------------
struct QuadPlan {
    bool ok;
    int tiles, nz, tpw, nwarps, nchunks, per_chunk, XN;
    size_t smem_bytes;
};
constexpr int kQuadTC = 4, kQuadTKC = 2, kQuadMaxWarps = 12;       // 384 threads: up to 168 registers each

static QuadPlan make_quad_plan(const Geometry &g)
{
    QuadPlan q{};
    q.ok = false;
    if (g.KH != 3 || g.KW != 3 || g.W % 4 != 0 || g.Cg % kQuadTC != 0 || g.Cg % kQuadTKC != 0 || g.Cg < 12 || g.B < 1)
        return q;
    q.tiles = (g.Cg / kQuadTC) * (g.Cg / kQuadTKC);
    // warps per CTA x tiles per warp: the combination that leaves the fewest idle tile slots
    q.tpw = 2;
    {
        int best_waste = 1 << 30;
        for (int nw : {12, 9, 8}) {
            const int nz = (q.tiles + nw * q.tpw - 1) / (nw * q.tpw);
            const int waste = nz * nw * q.tpw - q.tiles;
            if (waste < best_waste) { best_waste = waste; q.nwarps = nw; q.nz = nz; }
        }
        if (env().dw_cfg[0] > 0 && env().dw_cfg[0] <= kQuadMaxWarps) {       // tuning: warps per CTA
            q.nwarps = env().dw_cfg[0];
            q.tpw = (q.tiles + q.nwarps - 1) / q.nwarps;
            if (q.tpw > 4) q.tpw = 4;
            q.nz = (q.tiles + q.nwarps * q.tpw - 1) / (q.nwarps * q.tpw);
        }
    }
    // batch chunks: about a third of the SMs in total, but at least ~3 quads per lane and what fits in smem
    const int qpi = g.H * g.W / 4;
    q.XN = g.Cg * g.H * g.W;
    const size_t per_img = (size_t)2 * q.XN * sizeof(float);
    const int ctas = env().dw_cfg[1] > 0 ? env().dw_cfg[1] : kNumSM / 3 + 1;       // 50: two images per CTA at a batch of 100
    int want = (ctas + g.groups * q.nz - 1) / (g.groups * q.nz);
    if (want < 1) want = 1;
    int per_chunk = (g.B + want - 1) / want;
    while (per_chunk * qpi < 96 && per_chunk < g.B) per_chunk++;
    const int fit = (int)(((size_t)kMaxSmemBytes - 64) / per_img);
    if (fit < 1) return q;
    if (per_chunk > fit) per_chunk = fit;
    q.per_chunk = per_chunk;
    q.nchunks = (g.B + per_chunk - 1) / per_chunk;
    q.smem_bytes = 64 + (size_t)per_chunk * per_img;
    const size_t img_bytes = (size_t)q.XN * sizeof(float);
    if (img_bytes % 16 != 0) return q;
    q.ok = true;
    return q;
}

struct QuadParams {
    const float *dx, *y;
    float *partial;
    int B, C, H, W, Cg, ntk, tiles, tpw, per_chunk, XN, orient, bulk;
    int qpi, qpr;              // quads per image / per row
    unsigned m_qpi, m_qpr;     // ceil(2^32 / qpi), ceil(2^32 / qpr)
};

__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int KH, int KW, int TC, int TKC, bool FW>
__global__ void __launch_bounds__(kQuadMaxWarps * 32)
bwd_weight_quad_kernel(const QuadParams p)
{
    constexpr int K = KH * KW, NACC = TC * TKC * K;
    extern __shared__ __align__(128) float smem[];
    const int W = p.W, HW = p.H * p.W, Cg = p.Cg;
    const int chunk = blockIdx.x, G = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float *buf = smem + 16;                                   // [image][dX planes | y planes]
    const float *zero4 = smem + 4;                            // 16 zero bytes (see the neighbour loads)
    if (tid < 4) smem[4 + tid] = 0.f;
    const int b_begin = chunk * p.per_chunk;
    const int n_img = (b_begin + p.per_chunk < p.B ? b_begin + p.per_chunk : p.B) - b_begin;
    const size_t img_stride = (size_t)p.C * HW;
    const float *dx0 = p.dx + (size_t)G * Cg * HW + (size_t)b_begin * img_stride;
    const float *y0 = p.y + (size_t)G * Cg * HW + (size_t)b_begin * img_stride;
    const uint32_t img_bytes = (uint32_t)(Cg * HW) * 4u;
    const int XN = p.XN;

    if (p.bulk) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, 2u * img_bytes * (uint32_t)n_img);
            for (int i = 0; i < n_img; i++) {
                bulk_load(buf + (size_t)(2 * i) * XN, dx0 + (size_t)i * img_stride, img_bytes, bar);
                bulk_load(buf + (size_t)(2 * i + 1) * XN, y0 + (size_t)i * img_stride, img_bytes, bar);
            }
        }
        __syncthreads();
        mbar_wait(bar, 0);
    } else {
        for (int i = 0; i < n_img; i++)
            for (int e = tid; e < Cg * HW; e += blockDim.x) {
                buf[(size_t)(2 * i) * XN + e] = __ldg(dx0 + (size_t)i * img_stride + e);
                buf[(size_t)(2 * i + 1) * XN + e] = __ldg(y0 + (size_t)i * img_stride + e);
            }
        __syncthreads();
    }

    const bool fh = p.orient & 2;
    const int nquads = n_img * p.qpi;
    float *out = p.partial + ((size_t)chunk * p.C + (size_t)G * Cg) * Cg * K;   // [c][kc][t]
    // 32-bit shared-space addressing throughout the quad loop: generic 64-bit pointer arithmetic and pointer selects
    // were as many instructions as the FMAs themselves (266 of 554 per quad; every one costs an issue slot)
    const uint32_t buf_s = smem_u32(buf), zero_s = smem_u32(zero4);
    const uint32_t HW4 = (uint32_t)HW * 4u, img_b = (uint32_t)(2 * XN) * 4u, y_b = (uint32_t)XN * 4u;
    const uint32_t rstep4 = fh ? (uint32_t)W * 4u : (uint32_t)(-W * 4);       // row of tap qh: h - qh (h + qh reflected)
    const uint32_t side4 = FW ? 16u : (uint32_t)-16;

    for (int tw = 0; tw < p.tpw; tw++) {
        const int tile = (blockIdx.z * (blockDim.x >> 5) + warp) * p.tpw + tw;
        if (tile >= p.tiles) break;                              // warp-uniform
        const int c0 = (tile / p.ntk) * TC, k0 = (tile % p.ntk) * TKC;
        const uint32_t d_off = (uint32_t)(c0 * HW) * 4u, y_off = y_b + (uint32_t)(k0 * HW) * 4u;
        float acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = 0.f;
        for (int u = lane; u < nquads; u += 32) {
            const int img = p.m_qpi ? (int)__umulhi((unsigned)u, p.m_qpi) : u, r = u - img * p.qpi;     // magic 0: divisor 1
            const int h = p.m_qpr ? (int)__umulhi((unsigned)r, p.m_qpr) : r, wq = r - h * p.qpr;
            const uint32_t qbase = buf_s + (uint32_t)img * img_b + (uint32_t)(h * W + 4 * wq) * 4u;   // dX plane 0, the quad
            float4 d4[TC];
#pragma unroll
            for (int i = 0; i < TC; i++) d4[i] = lds_f4(qbase + d_off + (uint32_t)i * HW4);
            // window of 8 columns per (y channel, tap row): [4wq-4, 4wq+3], reflected: [4wq, 4wq+7]
            // neighbours outside the image read a zeroed 16-byte cell instead: no branch, no predicated load
            const bool side_ok = FW ? (wq + 1 < p.qpr) : (wq > 0);
            const uint32_t ybase = qbase + y_off;
#pragma unroll
            for (int qh = 0; qh < KH; qh++) {
                const int hh = fh ? h + qh : h - qh;
                const bool row_ok = (unsigned)hh < (unsigned)p.H;
                const uint32_t rowa = ybase + (uint32_t)qh * rstep4;
#pragma unroll
                for (int k = 0; k < TKC; k++) {
                    // a4: the quad's own columns, b4: the four columns beside it (channel k0 + k: k planes further)
                    const float4 a4 = lds_f4(row_ok ? rowa + (uint32_t)k * HW4 : zero_s);
                    const float4 b4 = lds_f4(row_ok && side_ok ? rowa + (uint32_t)k * HW4 + side4 : zero_s);
                    // win[j]: column 4wq - 4 + j (plain) / 4wq + j (reflected)
                    const float win[8] = {FW ? a4.x : b4.x, FW ? a4.y : b4.y, FW ? a4.z : b4.z, FW ? a4.w : b4.w,
                                          FW ? b4.x : a4.x, FW ? b4.y : a4.y, FW ? b4.z : a4.z, FW ? b4.w : a4.w};
                    // pixel j outermost: the TC*KW FMAs of one j are independent of each other, so the
                    // dependent chain of an accumulator is spaced TC*KW instructions apart
#pragma unroll
                    for (int j = 0; j < 4; j++) {
#pragma unroll
                        for (int qw = 0; qw < KW; qw++) {
#pragma unroll
                            for (int i = 0; i < TC; i++) {
                                const float dv = j == 0 ? d4[i].x : (j == 1 ? d4[i].y : (j == 2 ? d4[i].z : d4[i].w));
                                float &a = acc[((i * TKC + k) * KH + qh) * KW + qw];
                                a = fmaf(dv, win[FW ? j + qw : 4 + j - qw], a);
                            }
                        }
                    }
                }
            }
        }
        Halve<NACC, 32>::run(acc, lane);
        int off, size;
        rs_owner(NACC, 32, lane, &off, &size);
        constexpr int kFinal = halve_final(NACC, 32);
#pragma unroll
        for (int i = 0; i < kFinal; i++) {
            const int e = off + i;
            const int t = e % K, ik = (e / K) % TKC, ic = e / (K * TKC);
            if (i < size) out[((size_t)(c0 + ic) * Cg + (k0 + ik)) * K + t] = acc[i];
        }
    }
}

static int launch_bwd_weight_quad(const Geometry &g, const QuadPlan &q, const float *dx, const float *y, void *workspace,
                                  cudaStream_t s)
{
    QuadParams p{};
    p.dx = dx; p.y = y; p.partial = (float *)workspace;
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W; p.Cg = g.Cg;
    p.ntk = g.Cg / kQuadTKC; p.tiles = q.tiles; p.tpw = q.tpw; p.per_chunk = q.per_chunk; p.XN = q.XN;
    p.orient = g.orient;
    p.qpr = g.W / 4; p.qpi = g.H * p.qpr;
    // ceil(2^32 / d): u / d == umulhi(u, m) for the small indices used; d == 1 is flagged by m == 0
    p.m_qpi = p.qpi > 1 ? (unsigned)((0x100000000ULL + (unsigned)p.qpi - 1) / (unsigned)p.qpi) : 0u;
    p.m_qpr = p.qpr > 1 ? (unsigned)((0x100000000ULL + (unsigned)p.qpr - 1) / (unsigned)p.qpr) : 0u;
    p.bulk = (((uintptr_t)dx | (uintptr_t)y) % 16 == 0) && !env().nobulk ? 1 : 0;
    dim3 grid(q.nchunks, g.groups, q.nz);
    void (*kern)(const QuadParams) = (g.orient & 1) ? bwd_weight_quad_kernel<3, 3, kQuadTC, kQuadTKC, true>
                                                    : bwd_weight_quad_kernel<3, 3, kQuadTC, kQuadTKC, false>;
    if (q.smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.smem_bytes);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<grid, q.nwarps * 32, q.smem_bytes, s>>>(p);
    return cuda_status(cudaGetLastError());
}

// one WARP per dW element: lanes stride over the chunks, then a fixed butterfly -- the summation
// tree depends only on (nchunks), never on timing, so the result stays bit-reproducible
__global__ void __launch_bounds__(256)
bwd_weight_reduce_kernel(const float *__restrict__ partial, float *__restrict__ dw, int nchunks,
                         int C, int Cg, int Cw, int KH, int KW, size_t partial_stride, size_t dw_stride)
{
    partial += (size_t)blockIdx.y * partial_stride;        // batched: blockIdx.y = layer
    dw += (size_t)blockIdx.y * dw_stride;
    const int K = KH * KW;
    const int total = C * Cw * K;
    const size_t chunk_stride = (size_t)C * Cg * K;
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
    for (int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < total; e += warps_per_grid) {
        const int a = e % K;
        const int kc = (e / K) % Cw;
        const int c = e / (K * Cw);
        const int ah = a / KW, aw = a - ah * KW;
        const int qh = KH - 1 - ah, qw = KW - 1 - aw;
        const int cl = c % Cg;
        float s = 0.f;
        if (kc < Cg && !(qh == 0 && qw == 0 && kc >= cl)) {           // warp-uniform
            const float *src = partial + ((size_t)c * Cg + kc) * K + (qh * KW + qw);
            for (int n = lane; n < nchunks; n += 32) s += __ldg(src + (size_t)n * chunk_stride);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            s = -s;
        }
        if (lane == 0) dw[e] = s;
    }
}

// few chunks: one THREAD per dW element (a warp per element would leave most lanes idle)
__global__ void __launch_bounds__(256)
bwd_weight_reduce_thread_kernel(const float *__restrict__ partial, float *__restrict__ dw, int nchunks,
                                int C, int Cg, int Cw, int KH, int KW, size_t partial_stride, size_t dw_stride)
{
    partial += (size_t)blockIdx.y * partial_stride;
    dw += (size_t)blockIdx.y * dw_stride;
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
            // four independent partial sums in a FIXED association: deterministic, and the loads pipeline
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int n = 0;
            for (; n + 3 < nchunks; n += 4) {
                s0 += __ldg(src + (size_t)n * chunk_stride);
                s1 += __ldg(src + (size_t)(n + 1) * chunk_stride);
                s2 += __ldg(src + (size_t)(n + 2) * chunk_stride);
                s3 += __ldg(src + (size_t)(n + 3) * chunk_stride);
            }
            for (; n < nchunks; n++) s0 += __ldg(src + (size_t)n * chunk_stride);
            s = -((s0 + s1) + (s2 + s3));
        }
        dw[e] = s;
    }
}

static bool use_quad(const Geometry &g, QuadPlan *q)
{
    *q = make_quad_plan(g);
    return q->ok && !env().dw_quad_off;
}

static int stage1_chunks(const Geometry &g)
{
    QuadPlan q;
    if (use_quad(g, &q)) return q.nchunks;
    return make_plan(g).nchunks;
}

int launch_bwd_weight_partial(const Geometry &g, const float *dx, const float *y, void *workspace, cudaStream_t s)
{
    if (g.B == 0) return 0;
    QuadPlan q;
    if (use_quad(g, &q)) return launch_bwd_weight_quad(g, q, dx, y, workspace, s);
    const BwdWeightPlan pl = make_plan(g);
    BwdWeightParams p{};
    p.dx = dx; p.y = y; p.partial = (float *)workspace;
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W; p.KH = g.KH; p.KW = g.KW; p.Cg = g.Cg;
    p.ntk = pl.ntk; p.items = pl.items; p.per_chunk = pl.per_chunk;
    p.nbuf = pl.nbuf; p.nstage = pl.nstage; p.XN = pl.XN; p.orient = g.orient;
    p.w_magic = g.W > 1 ? (unsigned)((0x100000000ULL + g.W - 1) / g.W) : 0u;      // 0: single column
    const size_t img_bytes = (size_t)g.Cg * g.H * g.W * sizeof(float);
    p.bulk = (img_bytes % 16 == 0) && (((uintptr_t)dx | (uintptr_t)y) % 16 == 0) ? 1 : 0;
    if (env().nobulk) p.bulk = 0;
    dim3 grid(pl.nchunks, g.groups, pl.nz);
    void (*kern)(const BwdWeightParams) = nullptr;
    if (pl.tc == 4 && pl.tk == 12) kern = bwd_weight_partial_kernel<4, 12>;
    else if (pl.tc == 4 && pl.tk == 8) kern = bwd_weight_partial_kernel<4, 8>;
    else if (pl.tc == 4 && pl.tk == 4) kern = bwd_weight_partial_kernel<4, 4>;
    else if (pl.tc == 2) kern = bwd_weight_partial_kernel<2, 2>;
    else kern = bwd_weight_partial_kernel<1, 1>;
    if (pl.smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<grid, kWarps * 32, pl.smem_bytes, s>>>(p);
    return cuda_status(cudaGetLastError());
}

int launch_bwd_weight_reduce(const Geometry &g, int count, const void *workspace, size_t workspace_stride,
                             float *dw, size_t dw_stride, cudaStream_t s)
{
    if (count <= 0) return 0;
    const int total = g.C * g.Cw * g.K;
    const int nchunks = g.B > 0 ? stage1_chunks(g) : 0;
    const size_t pstride = workspace_stride / sizeof(float);
    if (nchunks >= 96) {                          // (a warp per element pays only for long chunk lists)
        int blocks = (total + 7) / 8;             // 8 warps per CTA, one element per warp
        const int cap = kNumSM * 8 / (count < 8 ? count : 8) + 1;
        if (blocks > cap) blocks = cap;
        dim3 grid(blocks, count);
        bwd_weight_reduce_kernel<<<grid, 256, 0, s>>>((const float *)workspace, dw, nchunks, g.C, g.Cg, g.Cw,
                                                      g.KH, g.KW, pstride, dw_stride);
    } else {
        int blocks = (total + 255) / 256;
        if (blocks > kNumSM * 8) blocks = kNumSM * 8;
        dim3 grid(blocks, count);
        bwd_weight_reduce_thread_kernel<<<grid, 256, 0, s>>>((const float *)workspace, dw, nchunks, g.C, g.Cg,
                                                             g.Cw, g.KH, g.KW, pstride, dw_stride);
    }
    return cuda_status(cudaGetLastError());
}

int launch_bwd_weight(const Geometry &g, const float *dx, const float *y, float *dw,
                      void *workspace, cudaStream_t s)
{
    int st = launch_bwd_weight_partial(g, dx, y, workspace, s);
    if (st != 0) return st;
    return launch_bwd_weight_reduce(g, 1, workspace, 0, dw, 0, s);
}

}  // namespace ifk
