// Wavefront triangular solve  out = L^-1 in  (reverse == false)  or  out = L^-T in
// (reverse == true, run as the same causal solve on the spatially reflected image).
//
// Replaces the reference's per-diagonal launch loop
// (inf/utils/inv_conv_cuda/inv_conv_with_bp_kernel_general.cu:72-129: (H+W-1)*C/4 launches,
// each followed by cudaDeviceSynchronize, one thread per (batch, group, pixel), dependent
// global read-modify-writes) by ONE launch: a CTA owns a (batch-stripe x channel-group)
// tile, keeps the image and the in-flight diagonals in shared memory, walks all H+W-1
// anti-diagonals internally with a block barrier per diagonal, holds its slice of the
// prepared k x k kernel in registers for the whole stripe, and reduces the Cg*k^2 receptive
// field with warp shuffles.
//
// Two kernels:
//   solve_smem_kernel<CC,CHUNK>  image (x and y, halo padded) resident in shared memory
//   solve_global_kernel          any shape; neighbours re-read from the output tensor (L1/L2)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ifk_internal.cuh"

namespace ifk {

struct SolveParams {
    const float *in;
    float *out;
    const float *prep;  // prepared weights of this direction: [group][co][KDP]
    int B, C, H, W, KH, KW, Cg, KD, KDP;
    int WP, CS;         // halo-padded row stride and channel stride (floats) of the y buffer
    int XN;             // floats per contiguous image buffer (Cg*H*W rounded up to 4)
    int NS, NCT, nslots;
    int reverse;
    int bulk;           // image size / pointers allow TMA bulk copies (16-byte granularity)
};

// ---- TMA bulk copy / mbarrier primitives (PTX; SASS: UBLKCP, SYNCS) -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "IFK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra IFK_DONE_%=;\n\t"
        "bra IFK_WAIT_%=;\n\t"
        "IFK_DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_proxy()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// Shared-memory resident kernel.
//
// Per image of the CTA's batch stripe:
//   1. the group's image (contiguous in NCHW) lands in `xbuf` by one TMA bulk copy;
//   2. pre-pass, no dependencies: z = T x for every pixel into `zbuf` (skipped when Cg == 1);
//      xbuf is then free and the NEXT image of the stripe is prefetched into it;
//   3. wavefront: thread -> (row slot, ct, ks) walks its image row, one pixel per
//      anti-diagonal; `ct` = tile of CC output channels, `ks` = slice of the (K-1)*Cg
//      neighbour reduction (kidx = j*NS + ks, j < CHUNK) whose weights stay in registers;
//      partial sums are combined with warp shuffles, the ks == 0 lane adds z and writes y
//      both into the halo-padded `ybuf` (read by later diagonals) and in place into `zbuf`;
//      one block barrier per diagonal;
//   4. `zbuf` goes back to global memory by one TMA bulk store.
// The adjoint solve (reverse) is the same walk in reflected coordinates: only the index into
// the contiguous buffers is mirrored.
// ------------------------------------------------------------------------------------------
template <int CC, int CHUNK>
constexpr int solve_max_threads()
{
    // registers: CC*CHUNK weights + CHUNK offsets + ~56 of bookkeeping
    int regs = CC * CHUNK + CHUNK + 56;
    if (regs > 255) regs = 255;
    int t = (65536 / regs) / 32 * 32;
    return t > 1024 ? 1024 : t;
}

template <int CC, int CHUNK>
__global__ void __launch_bounds__(solve_max_threads<CC, CHUNK>())
solve_smem_kernel(const SolveParams p)
{
    extern __shared__ __align__(128) float smem[];
    const int Cg = p.Cg, CS = p.CS, WP = p.WP, H = p.H, W = p.W, HW = p.H * p.W;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);     // 16 bytes reserved
    float *xbuf = smem + 4;                                  // [Cg][HW] raw input
    float *zbuf = Cg > 1 ? xbuf + p.XN : xbuf;               // [Cg][HW] T x, then y in place
    float *ybuf = zbuf + p.XN;                               // [Cg][CS] y with zero halo (top/left)

    const int tid = threadIdx.x;
    const int NS = p.NS, NCT = p.NCT;
    const int ks = tid % NS;
    const int ct = (tid / NS) % NCT;
    const int slot = tid / (NS * NCT);
    const bool worker = slot < p.nslots;
    const int G = blockIdx.y;
    const unsigned lane = tid & 31u;
    const unsigned gmask = NS >= 32 ? 0xffffffffu : (((1u << NS) - 1u) << (lane & ~(unsigned)(NS - 1)));
    const float *wg = p.prep + (size_t)G * Cg * p.KDP;
    const uint32_t img_bytes = (uint32_t)(Cg * HW) * 4u;
    const size_t img_stride = (size_t)p.C * HW;
    const float *in0 = p.in + (size_t)G * Cg * HW;
    float *out0 = p.out + (size_t)G * Cg * HW;

    int b = blockIdx.x;
    if (p.bulk && tid == 0) {
        mbar_init(bar, 1);
        if (b < p.B) {
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b * img_stride, img_bytes, bar);
        }
    }

    // this thread's slice of the prepared kernel -> registers, for the whole batch stripe
    float wreg[CC][CHUNK];
    int offs[CHUNK];
    {
        const int KDY = p.KD - Cg;     // neighbour taps only; tap 0 (T) is applied in the pre-pass
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {
            const int kidx = j * NS + ks;
            const bool valid = worker && kidx < KDY;
            const int t = valid ? 1 + kidx / Cg : 1;
            const int ci = valid ? kidx - (t - 1) * Cg : 0;
            const int qh = t / p.KW, qw = t - qh * p.KW;
            offs[j] = valid ? ci * CS - qh * WP - qw : 0;   // padding entries: weight 0, finite data
#pragma unroll
            for (int cc = 0; cc < CC; cc++) {
                const int co = ct * CC + cc;
                wreg[cc][j] = (valid && co < Cg) ? __ldg(wg + (size_t)co * p.KDP + Cg + kidx) : 0.f;
            }
        }
    }

    __syncthreads();                   // mbarrier initialised before anyone waits on it

    const int ndiag = H + W - 1;
    const int halo = (p.KH - 1) * WP + (p.KW - 1);
    uint32_t parity = 0;
    for (; b < p.B; b += gridDim.x) {
        const int b_next = b + gridDim.x;
        // ybuf starts from zero for every image (halo, and the not-yet-written interior that
        // zero-weight padding entries may touch)
        for (int i = tid; i < Cg * CS; i += blockDim.x) ybuf[i] = 0.f;
        if (p.bulk) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            const float *src = in0 + (size_t)b * img_stride;
            for (int i = tid; i < Cg * HW; i += blockDim.x) xbuf[i] = __ldg(src + i);
        }
        if (Cg > 1) {
            if (tid == 0 && p.bulk) bulk_store_wait_read();     // previous image has left zbuf
            __syncthreads();
            // pre-pass z = T x : item (co, r), consecutive threads -> consecutive pixels
            for (int i = tid; i < Cg * HW; i += blockDim.x) {
                const int co = i / HW, r = i - co * HW;
                const float *tr = wg + (size_t)co * p.KDP;
                float a0 = 0.f, a1 = 0.f;
                int ci = 0;
                for (; ci + 1 < Cg; ci += 2) {
                    a0 = fmaf(__ldg(tr + ci), xbuf[ci * HW + r], a0);
                    a1 = fmaf(__ldg(tr + ci + 1), xbuf[(ci + 1) * HW + r], a1);
                }
                if (ci < Cg) a0 = fmaf(__ldg(tr + ci), xbuf[ci * HW + r], a0);
                zbuf[i] = a0 + a1;
            }
        }
        __syncthreads();
        if (p.bulk && tid == 0 && Cg > 1 && b_next < p.B) {        // prefetch the next image
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b_next * img_stride, img_bytes, bar);
        }

        for (int d = 0; d < ndiag; d++) {
            if (worker) {
                for (int h = slot; h < H; h += p.nslots) {
                    const int w = d - h;
                    if (w < 0 || w >= W) continue;
                    const int rr = h * W + w;
                    const int r = p.reverse ? HW - 1 - rr : rr;
                    float *zp = zbuf + (size_t)(ct * CC) * HW + r;
                    float zv[CC];
#pragma unroll
                    for (int cc = 0; cc < CC; cc++)
                        zv[cc] = (ks == 0 && ct * CC + cc < Cg) ? zp[cc * HW] : 0.f;
                    const float *px = ybuf + h * WP + w + halo;
                    float acc0[CC], acc1[CC];
#pragma unroll
                    for (int cc = 0; cc < CC; cc++) acc0[cc] = acc1[cc] = 0.f;
#pragma unroll
                    for (int j = 0; j < CHUNK; j++) {
                        const float v = px[offs[j]];
#pragma unroll
                        for (int cc = 0; cc < CC; cc++) {
                            if (j & 1) acc1[cc] = fmaf(wreg[cc][j], v, acc1[cc]);
                            else       acc0[cc] = fmaf(wreg[cc][j], v, acc0[cc]);
                        }
                    }
#pragma unroll
                    for (int cc = 0; cc < CC; cc++) acc0[cc] += acc1[cc];
                    for (int m = NS >> 1; m > 0; m >>= 1) {
#pragma unroll
                        for (int cc = 0; cc < CC; cc++)
                            acc0[cc] += __shfl_xor_sync(gmask, acc0[cc], m);
                    }
                    if (ks == 0) {
                        float *py = ybuf + h * WP + w + halo + ct * CC * CS;
#pragma unroll
                        for (int cc = 0; cc < CC; cc++)
                            if (ct * CC + cc < Cg) {
                                const float yv = acc0[cc] + zv[cc];
                                py[cc * CS] = yv;
                                zp[cc * HW] = yv;
                            }
                    }
                }
            }
            if (blockDim.x <= 32) __syncwarp(); else __syncthreads();
        }

        float *dst = out0 + (size_t)b * img_stride;
        if (p.bulk) {
            fence_async_proxy();            // generic-proxy writes of zbuf -> visible to the TMA engine
            __syncthreads();
            if (tid == 0) {
                bulk_store(dst, zbuf, img_bytes);
                if (Cg == 1) {              // zbuf aliases xbuf: reuse only after the store has read it
                    bulk_store_wait_read();
                    if (b_next < p.B) {
                        mbar_expect_tx(bar, img_bytes);
                        bulk_load(xbuf, in0 + (size_t)b_next * img_stride, img_bytes, bar);
                    }
                }
            }
        } else {
            __syncthreads();
            for (int i = tid; i < Cg * HW; i += blockDim.x) dst[i] = zbuf[i];
            __syncthreads();
        }
    }
    if (p.bulk && tid == 0) bulk_store_wait_read();   // smem must outlive the last store's read
}

// ------------------------------------------------------------------------------------------
// Fallback for images that do not fit in shared memory: same algorithm, neighbours read
// back from the output tensor (visible to the whole CTA after the per-diagonal barrier).
// One thread per (pixel on the diagonal, output channel).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
solve_global_kernel(const SolveParams p)
{
    const int Cg = p.Cg, H = p.H, W = p.W, HW = p.H * p.W, KW = p.KW, K = p.KH * p.KW;
    const int G = blockIdx.y;
    const float *wg = p.prep + (size_t)G * Cg * p.KDP;
    const int ndiag = H + W - 1;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const size_t gbase = ((size_t)b * p.C + (size_t)G * Cg) * HW;
        const float *in = p.in + gbase;
        float *out = p.out + gbase;
        for (int d = 0; d < ndiag; d++) {
            const int hmin = d - (W - 1) > 0 ? d - (W - 1) : 0;
            const int hmax = d < H - 1 ? d : H - 1;
            const int work = (hmax - hmin + 1) * Cg;
            for (int e = threadIdx.x; e < work; e += blockDim.x) {
                const int co = e % Cg, h = hmin + e / Cg, w = d - h;
                const float *wr = wg + (size_t)co * p.KDP;
                const int r = h * W + w;
                const int gr = p.reverse ? HW - 1 - r : r;
                float acc = 0.f;
                for (int ci = 0; ci < Cg; ci++) acc = fmaf(__ldg(wr + ci), in[ci * HW + gr], acc);
                for (int t = 1; t < K; t++) {
                    const int qh = t / KW, qw = t - qh * KW;
                    if (h - qh < 0 || w - qw < 0) continue;
                    const int rn = (h - qh) * W + (w - qw);
                    const int gn = p.reverse ? HW - 1 - rn : rn;
                    const float *wt = wr + t * Cg;
                    for (int ci = 0; ci < Cg; ci++)
                        acc = fmaf(__ldg(wt + ci), __ldcg(out + ci * HW + gn), acc);
                }
                __stcg(out + co * HW + gr, acc);    // in and out never alias (ifk.h)
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Host side: variant selection and launch.
// ------------------------------------------------------------------------------------------
struct SolveConfig {
    bool smem;      // false -> global fallback
    int cc, chunk, ns, nct, nslots, threads;
    int WP, CS, XN;
    size_t smem_bytes;
    int grid_x;
};

static const int kCCs[] = {1, 2, 3, 4};
static const int kChunks[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18, 20, 24, 28, 32};

static int max_threads_for(int cc, int chunk)
{
    int regs = cc * chunk + chunk + 56;
    if (regs > 255) regs = 255;
    int t = (65536 / regs) / 32 * 32;
    return t > 1024 ? 1024 : t;
}

static int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) l++; return l; }

static bool parse_forced(int *cc, int *chunk, int *ns, int *nslots)
{
    const char *e = getenv("IFK_SOLVE_CFG");   // "cc,chunk,ns,nslots" -- tuning experiments only
    if (!e || !*e) return false;
    return sscanf(e, "%d,%d,%d,%d", cc, chunk, ns, nslots) == 4;
}

static SolveConfig choose_config(const Geometry &g)
{
    SolveConfig best{};
    best.smem = false;
    best.threads = 512;
    best.grid_x = g.B < 4 * kNumSM ? g.B : 4 * kNumSM;
    if (best.grid_x < 1) best.grid_x = 1;

    const int HP = g.H + g.KH - 1, WP = g.W + g.KW - 1;
    int CS = HP * WP;
    if ((CS & 1) == 0) CS += 1;   // odd channel stride: channels of one pixel hit distinct banks
    const int XN = round_up(g.Cg * g.H * g.W, 4);
    const int nx = g.Cg > 1 ? 2 : 1;
    const size_t smem_bytes = 16 + ((size_t)nx * XN + (size_t)g.Cg * CS) * sizeof(float);
    const char *force_global = getenv("IFK_SOLVE_GLOBAL");
    if (smem_bytes > (size_t)kMaxSmemBytes || (force_global && force_global[0] == '1')) return best;

    const int KDY = g.KD - g.Cg;            // neighbour taps (the centre tap is the pre-pass)
    const int rows = g.H;
    double best_cost = 1e30;
    int fcc, fchunk, fns, fslots;
    const bool forced = parse_forced(&fcc, &fchunk, &fns, &fslots);
    for (int cc : kCCs) {
        if (cc > g.Cg) continue;
        const int nct = (g.Cg + cc - 1) / cc;
        for (int chunk : kChunks) {
            for (int ns = 1; ns <= 32; ns *= 2) {
                if ((long)ns * chunk < KDY) continue;
                if (ns > 1 && (long)(ns / 2) * chunk >= KDY) continue;      // needless split
                if (chunk > 1 && (long)ns * (chunk - 1) >= KDY && KDY > 0) {
                    bool smaller_listed = false;                            // a tighter chunk exists
                    for (int c2 : kChunks) smaller_listed |= (c2 < chunk && (long)ns * c2 >= KDY);
                    if (smaller_listed) continue;
                }
                const int per_slot = ns * nct;
                const int tmax = max_threads_for(cc, chunk);
                if (per_slot > tmax) continue;
                int nslots = tmax / per_slot;
                if (nslots > rows) nslots = rows;
                if (forced) {
                    if (cc != fcc || chunk != fchunk || ns != fns) continue;
                    if (fslots > 0 && fslots <= nslots) nslots = fslots;
                }
                for (; nslots >= 1; nslots = forced ? 0 : nslots / 2) {
                    const int threads = round_up(nslots * per_slot, 32);
                    const int iters = (rows + nslots - 1) / nslots;
                    const int warps = threads / 32;
                    const double instr = chunk * (2.0 + cc) + 2.0 * cc * ilog2(ns) + 30.0;
                    const double waste = (double)(nct * cc) / g.Cg;
                    const double issue = instr * ((warps + 3) / 4);
                    const double latency = instr + 60.0 + 25.0 * ilog2(ns);
                    const double barrier = warps > 1 ? 20.0 + 2.0 * warps : 5.0;
                    double cost = iters * (issue > latency ? issue : latency) + barrier;
                    cost *= 1.0 + 0.05 * (waste - 1.0);
                    if (cost < best_cost) {
                        best_cost = cost;
                        best.smem = true;
                        best.cc = cc; best.chunk = chunk; best.ns = ns; best.nct = nct;
                        best.nslots = nslots; best.threads = threads;
                    }
                    if (nslots == 1) break;
                }
            }
        }
    }
    if (!best.smem) return best;
    best.WP = WP; best.CS = CS; best.XN = XN; best.smem_bytes = smem_bytes;
    // CTAs resident per SM (shared memory and thread limits), then one stripe of images each
    int per_sm = (int)((size_t)(kMaxSmemBytes + 1024) / (smem_bytes + 1024));
    const int by_threads = 2048 / best.threads;
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int grid_x = (kNumSM * per_sm + g.groups - 1) / g.groups;
    if (grid_x > g.B) grid_x = g.B;
    if (grid_x < 1) grid_x = 1;
    best.grid_x = grid_x;
    return best;
}

template <int CC, int CHUNK>
static int launch_smem_variant(const SolveParams &p, const SolveConfig &c, int groups, cudaStream_t s)
{
    auto kern = solve_smem_kernel<CC, CHUNK>;
    if (c.smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)c.smem_bytes);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(c.grid_x, groups);
    kern<<<grid, c.threads, c.smem_bytes, s>>>(p);
    return cuda_status(cudaGetLastError());
}

template <int CC>
static int dispatch_chunk(const SolveParams &p, const SolveConfig &c, int groups, cudaStream_t s)
{
    switch (c.chunk) {
#define IFK_CASE(N) case N: return launch_smem_variant<CC, N>(p, c, groups, s);
        IFK_CASE(1) IFK_CASE(2) IFK_CASE(3) IFK_CASE(4) IFK_CASE(5) IFK_CASE(6) IFK_CASE(7) IFK_CASE(8)
        IFK_CASE(9) IFK_CASE(10) IFK_CASE(12) IFK_CASE(14) IFK_CASE(16) IFK_CASE(18) IFK_CASE(20)
        IFK_CASE(24) IFK_CASE(28) IFK_CASE(32)
#undef IFK_CASE
    }
    return IFK_ERR_UNSUPPORTED;
}

int launch_solve(const Geometry &g, const float *in, const float *prep_dir, float *out,
                 bool reverse, cudaStream_t s)
{
    if (g.B == 0) return 0;
    const SolveConfig c = choose_config(g);
    SolveParams p{};
    p.in = in; p.out = out; p.prep = prep_dir;
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W; p.KH = g.KH; p.KW = g.KW;
    p.Cg = g.Cg; p.KD = g.KD; p.KDP = g.KDP;
    p.reverse = reverse ? 1 : 0;
    if (!c.smem) {
        dim3 grid(c.grid_x, g.groups);
        solve_global_kernel<<<grid, c.threads, 0, s>>>(p);
        return cuda_status(cudaGetLastError());
    }
    p.WP = c.WP; p.CS = c.CS; p.XN = c.XN; p.NS = c.ns; p.NCT = c.nct; p.nslots = c.nslots;
    const size_t img_bytes = (size_t)g.Cg * g.H * g.W * sizeof(float);
    p.bulk = (img_bytes % 16 == 0) && (((uintptr_t)in | (uintptr_t)out) % 16 == 0) ? 1 : 0;
    if (const char *nb = getenv("IFK_SOLVE_NOBULK")) if (nb[0] == '1') p.bulk = 0;
    switch (c.cc) {
        case 1: return dispatch_chunk<1>(p, c, g.groups, s);
        case 2: return dispatch_chunk<2>(p, c, g.groups, s);
        case 3: return dispatch_chunk<3>(p, c, g.groups, s);
        case 4: return dispatch_chunk<4>(p, c, g.groups, s);
    }
    return IFK_ERR_UNSUPPORTED;
}

int describe_solve(const Geometry &g, char *buf, size_t buflen)
{
    const SolveConfig c = choose_config(g);
    if (c.smem)
        snprintf(buf, buflen, "smem<cc=%d,chunk=%d> ns=%d nct=%d slots=%d threads=%d smem=%zuB grid=%dx%d",
                 c.cc, c.chunk, c.ns, c.nct, c.nslots, c.threads, c.smem_bytes, c.grid_x, g.groups);
    else
        snprintf(buf, buflen, "global threads=%d grid=%dx%d", c.threads, c.grid_x, g.groups);
    return 0;
}

}  // namespace ifk
