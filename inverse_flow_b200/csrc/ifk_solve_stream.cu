// Wavefront solve for images that do not fit in shared memory ("stream" kernel).
//
// Same decomposition as the resident kernel (ifk_solve_kernel.cuh): thread = (image row slot,
// tile of CC output channels, slice `ks` of the (K-1)*Cg neighbour reduction), the slice's
// prepared weights in registers for the whole batch stripe, reduce-scatter over warp shuffles,
// one named barrier per anti-diagonal.  What changes is where the image lives: nothing is
// staged; the pre-pass writes z = T x straight into the OUTPUT tensor, the wavefront reads its
// neighbours back from the output tensor (L1/L2; the CTA barrier orders its own global writes
// at CTA scope) and overwrites z with y in place.  No workspace, any H x W; the only limit is
// that the weight slices fit the register file: Cg*Cg*(K-1) <= ~48K per CTA.
//
// Cluster mode (CL = true) lifts that limit: a thread-block cluster of 2..8 CTAs shares one
// image, each CTA owning a slice of the OUTPUT channels (and the registers for its weights);
// y is exchanged through the output tensor in L2 (ld.global.cg) and the per-diagonal barrier
// becomes a cluster barrier (barrier.cluster.arrive.release / wait.acquire).  This is what makes
// Cg = 96 (73.7K weights at k = 3, more than one SM's register file) run at all.
//
// Replaces, for large images, the same reference loop as the resident kernel
// (inv_conv_with_bp_kernel_general.cu:72-129).
#include <stdio.h>
#include <stdlib.h>
#include "ifk_env.cuh"
#include "ifk_solve_kernel.cuh"

namespace ifk {

struct StreamParams {
    const float *in;
    float *out;
    const float *prep;
    int B, C, H, W, KH, KW, Cg, KDP, CgP4;
    int NVT;            // (K-1)*Cg reduction entries per output
    int NS, NCT, nslots, iters, nwork;
    int kw_magic, v_dt, v_dq;
    int flip;           // as SolveParams::flip
    int csize;          // CTAs per cluster (1 = no cluster); CTA r owns channel tiles [r*NCT, (r+1)*NCT)
};

// reduce-scatter as Rs (ifk_solve_kernel.cuh), finishing into global memory
template <int N, int LEVELS>
struct RsGlobal {
    __device__ __forceinline__ static void run(float *acc, const float *zv, int ks, int m, int own_size,
                                               bool active, float *dst, int stride)
    {
        if (m == 0 || LEVELS == 0) {
#pragma unroll
            for (int i = 0; i < N; i++)
                if (active && i < own_size) dst[(size_t)i * stride] = acc[i] + zv[i];
            return;
        }
        constexpr int HALF = (N + 1) / 2;
        const bool hi = (ks & m) != 0;
#pragma unroll
        for (int i = 0; i < HALF; i++) {
            const float lo_v = acc[i];
            const float hi_v = (i + HALF < N) ? acc[i + HALF] : 0.f;
            acc[i] = (hi ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, hi ? lo_v : hi_v, m);
        }
        RsGlobal<HALF, (LEVELS > 0 ? LEVELS - 1 : 0)>::run(acc, zv, ks, m >> 1, own_size, active, dst, stride);
    }
};

template <int CC, int NV>
constexpr int stream_max_threads()
{
    int regs = CC * NV + 3 * NV + 2 * CC + 56;
    if (regs > 255) regs = 255;
    int t = (65536 / regs) / 32 * 32;
    return t > 1024 ? 1024 : t;
}

__device__ __forceinline__ float ld_cg(const float *p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

template <int CC, int NV, bool CL>
__global__ void __launch_bounds__(stream_max_threads<CC, NV>())
solve_stream_kernel(const StreamParams p)
{
    extern __shared__ __align__(16) float tT[];     // [Cg][CgP4] transposed T
    const int Cg = p.Cg, H = p.H, W = p.W, HW = p.H * p.W;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int G = blockIdx.y;
    const float *wg = p.prep + (size_t)G * Cg * p.KDP;

    const int NS = p.NS, NCT = p.NCT;
    const int ks = tid % NS;
    const int crank = CL ? (int)cluster_ctarank() : 0;
    const int ct = crank * NCT + (tid / NS) % NCT;          // global channel tile of this thread
    const int slot = tid / (NS * NCT);
    const bool worker = slot < p.nslots;

    float wreg[CC][NV];
    int offs[NV], qhw[NV];
    {
        const int sw = (p.flip & 1) ? -1 : 1, sh = (p.flip & 2) ? -1 : 1;   // reflected axis: neighbours lie ahead in memory
        int t1 = ks / Cg, q = ks - t1 * Cg;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const bool valid = worker && j * NS + ks < p.NVT;
            const int t = t1 + 1;
            const int qh = (t * p.kw_magic) >> 16, qw = t - qh * p.KW;
            offs[j] = valid ? q * HW - (sh * qh * W + sw * qw) : 0;
            qhw[j] = valid ? ((qh << 8) | qw) : 0x7f7f;      // padding entries never pass the border test
#pragma unroll
            for (int cc = 0; cc < CC; cc++) {
                const int co = ct * CC + cc;
                wreg[cc][j] = (valid && co < Cg) ? __ldg(wg + (size_t)co * p.KDP + Cg + t1 * Cg + q) : 0.f;
            }
            q += p.v_dq;
            t1 += p.v_dt;
            if (q >= Cg) { q -= Cg; t1++; }
        }
    }
    for (int i = tid; i < Cg * p.CgP4; i += nthr) {
        const int ci = i / p.CgP4, co = i - ci * p.CgP4;
        tT[i] = co < Cg ? __ldg(wg + (size_t)co * p.KDP + ci) : 0.f;
    }
    __syncthreads();

    int own_off, own_size;
    rs_owner(CC, NS, ks, &own_off, &own_size);
    {
        const int tile_n = Cg - ct * CC < CC ? Cg - ct * CC : CC;
        own_size = own_off + own_size > tile_n ? (tile_n - own_off > 0 ? tile_n - own_off : 0) : own_size;
    }
    const int own_c0 = ct * CC + own_off;
    const int ndiag = H + W - 1;
    const int KH1 = p.KH - 1, KW1 = p.KW - 1;
    const int sw_1 = (p.flip & 1) ? -1 : 1, sh_w = (p.flip & 2) ? -W : W;
    const int idx0 = ((p.flip & 2) ? (H - 1) * W : 0) + ((p.flip & 1) ? W - 1 : 0);

    const int nclusters = CL ? gridDim.x / p.csize : gridDim.x;
    for (int b = CL ? blockIdx.x / p.csize : blockIdx.x; b < p.B; b += nclusters) {
        const size_t gbase = ((size_t)b * p.C + (size_t)G * Cg) * HW;
        const float *in_b = p.in + gbase;
        float *out_b = p.out + gbase;

        // pre-pass z = T x, pointwise, coalesced over the pixels; z goes straight to the output
        const int n4 = p.CgP4 >> 2;
        for (int i = (CL ? crank * nthr : 0) + tid; i < HW * n4; i += (CL ? p.csize : 1) * nthr) {
            const int c4 = i / HW, r = i - c4 * HW;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const float *tp = tT + c4 * 4;
#pragma unroll 4
            for (int ci = 0; ci < Cg; ci++) {
                const float xv = __ldg(in_b + (size_t)ci * HW + r);
                const float4 t4 = *reinterpret_cast<const float4 *>(tp + ci * p.CgP4);
                a0 = fmaf(t4.x, xv, a0);
                a1 = fmaf(t4.y, xv, a1);
                a2 = fmaf(t4.z, xv, a2);
                a3 = fmaf(t4.w, xv, a3);
            }
            const int co = c4 * 4;
            out_b[(size_t)co * HW + r] = a0;
            if (co + 1 < Cg) out_b[(size_t)(co + 1) * HW + r] = a1;
            if (co + 2 < Cg) out_b[(size_t)(co + 2) * HW + r] = a2;
            if (co + 3 < Cg) out_b[(size_t)(co + 3) * HW + r] = a3;
        }
        if (CL) cluster_barrier();   // z visible to the whole cluster (release/acquire at cluster scope)
        else __syncthreads();        // z visible to the whole CTA (barrier = CTA-scope fence)

        if (CL || tid < p.nwork) {       // cluster barriers need every thread of every CTA
            for (int d = 0; d < ndiag; d++) {
                for (int it = 0; it < p.iters; it++) {
                    const int h = slot + it * p.nslots;
                    const int w = d - h;
                    const bool active = worker && h < H && (unsigned)w < (unsigned)W;
                    if (!__any_sync(0xffffffffu, active)) continue;        // warp-uniform
                    const int m = active ? idx0 + sh_w * h + sw_1 * w : 0;   // memory index of solver pixel (h, w)
                    float *base = out_b + m;
                    const int hh = active ? h : -1, ww = active ? w : -1;  // idle lanes load nothing

                    float v[NV];
                    if (hh >= KH1 && ww >= KW1) {                          // interior: no border tests
#pragma unroll
                        for (int j = 0; j < NV; j++) v[j] = CL ? ld_cg(base + offs[j]) : base[offs[j]];   // padding: offset 0, weight 0
                    } else {
#pragma unroll
                        for (int j = 0; j < NV; j++)
                            v[j] = (hh >= (qhw[j] >> 8) && ww >= (qhw[j] & 0xff)) ? (CL ? ld_cg(base + offs[j]) : base[offs[j]]) : 0.f;
                    }
                    float zv[CC];
#pragma unroll
                    for (int i = 0; i < CC; i++)
                        zv[i] = (active && i < own_size) ? (CL ? ld_cg(base + (size_t)(own_c0 + i) * HW) : base[(size_t)(own_c0 + i) * HW]) : 0.f;

                    float acc[CC];
#pragma unroll
                    for (int cc = 0; cc < CC; cc++) acc[cc] = 0.f;
#pragma unroll
                    for (int j = 0; j < NV; j++)
#pragma unroll
                        for (int cc = 0; cc < CC; cc++) acc[cc] = fmaf(wreg[cc][j], v[j], acc[cc]);
                    RsGlobal<CC, 5>::run(acc, zv, ks, NS >> 1, own_size, active, base + (size_t)own_c0 * HW, HW);
                }
                if (CL) cluster_barrier();
                else asm volatile("bar.sync 1, %0;" ::"r"(p.nwork) : "memory");    // worker warps; CTA-scope fence
            }
        }
        if (CL) cluster_barrier(); else __syncthreads();
    }
}

// ---- host side -----------------------------------------------------------------------------
struct StreamConfig {
    bool ok;
    int cc, nv, ns, nct, nslots, iters, threads, nwork, grid_x;
    int csize;          // CTAs per cluster sharing one image (output-channel split); nct is per CTA
    size_t smem_bytes;
};

#define IFK_STREAM_VARIANTS                                                                          \
    X(1, 8) X(1, 12) X(1, 24) X(2, 6) X(2, 12) X(2, 24) X(3, 6) X(3, 9) X(3, 12) X(3, 24)            \
    X(4, 3) X(4, 6) X(4, 9) X(4, 12) X(4, 18) X(4, 24) X(6, 3) X(6, 6) X(6, 9) X(6, 12) X(6, 18)     \
    X(8, 3) X(8, 6) X(8, 12) X(8, 18) X(12, 3) X(12, 6) X(12, 9) X(12, 12)

static int stream_variant_threads(int cc, int nv)
{
#define X(CC, NV) if (cc == CC && nv == NV) return stream_max_threads<CC, NV>();
    IFK_STREAM_VARIANTS
#undef X
    return 0;
}

static StreamConfig choose_stream(const Geometry &g)
{
    StreamConfig best{};
    best.ok = false;
    const int NVT = (g.K - 1) * g.Cg;
    const size_t smem = (size_t)g.Cg * round_up(g.Cg, 4) * sizeof(float);
    if (smem > (size_t)kMaxSmemBytes || NVT == 0) return best;
    double best_cost = 1e30;
    int fcc = 0, fnv = 0;
    fcc = env().stream_cfg[0]; fnv = env().stream_cfg[1];   // IFK_STREAM_CFG, tuning only
    static const int kCCs[] = {12, 8, 6, 4, 3, 2, 1};
    static const int kNVs[] = {3, 6, 8, 9, 12, 18, 24};
    static const int kCsizes[] = {1, 2, 4, 8};
    for (int csize : kCsizes) {
        if (best.ok) break;                      // smallest cluster that holds the weights wins
        for (int cc : kCCs) {
            if (cc > g.Cg) continue;
            const int nct_total = (g.Cg + cc - 1) / cc;
            if (csize > nct_total) continue;
            const int nct = (nct_total + csize - 1) / csize;
            for (int nv : kNVs) {
                const int tmax = stream_variant_threads(cc, nv);
                if (tmax == 0) continue;
                if (fcc && (cc != fcc || nv != fnv)) continue;
                for (int ns = 1; ns <= 32; ns *= 2) {
                    if ((long)ns * nv < NVT) continue;
                    if (ns > 1 && (long)(ns / 2) * nv >= NVT) continue;
                    const int per_slot = ns * nct;
                    if (per_slot > tmax) continue;
                    int nslots = tmax / per_slot;
                    if (nslots > g.H) nslots = g.H;
                    const int iters = (g.H + nslots - 1) / nslots;
                    const int threads = round_up(nslots * per_slot, 32);
                    // work per diagonal ~ iters * (loads + FMAs + shuffles), all slots in parallel
                    const double waste = (double)(ns * nv) / NVT * (double)(nct * csize * cc) / g.Cg;
                    const double cost = iters * (nv * (2.0 + cc) + 6.0 * cc + 60.0) * waste * ((threads + 127) / 128);
                    if (cost < best_cost) {
                        best_cost = cost;
                        best.ok = true;
                        best.cc = cc; best.nv = nv; best.ns = ns; best.nct = nct; best.nslots = nslots;
                        best.iters = iters; best.threads = threads; best.nwork = threads; best.csize = csize;
                    }
                }
            }
        }
    }
    if (!best.ok) return best;
    best.smem_bytes = smem;
    int per_sm = 2048 / best.threads;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int nclusters = (kNumSM * per_sm / best.csize + g.groups - 1) / g.groups;
    if (nclusters > g.B) nclusters = g.B;
    if (nclusters < 1) nclusters = 1;
    best.grid_x = nclusters * best.csize;
    return best;
}

bool stream_solve_available(const Geometry &g) { return choose_stream(g).ok; }

int describe_stream_solve(const Geometry &g, char *buf, size_t buflen)
{
    const StreamConfig c = choose_stream(g);
    snprintf(buf, buflen, "stream<cc=%d,nv=%d> cluster=%d ns=%d nct=%d slots=%d iters=%d threads=%d grid=%dx%d",
             c.cc, c.nv, c.csize, c.ns, c.nct, c.nslots, c.iters, c.threads, c.grid_x, g.groups);
    return 0;
}

int launch_solve_stream(const Geometry &g, const float *in, const float *prep_dir, float *out,
                        bool reverse, cudaStream_t s)
{
    const StreamConfig c = choose_stream(g);
    if (!c.ok) return IFK_ERR_UNSUPPORTED;
    StreamParams p{};
    p.in = in; p.out = out; p.prep = prep_dir;
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W; p.KH = g.KH; p.KW = g.KW; p.Cg = g.Cg; p.KDP = g.KDP;
    p.CgP4 = round_up(g.Cg, 4);
    p.NVT = (g.K - 1) * g.Cg;
    p.NS = c.ns; p.NCT = c.nct; p.nslots = c.nslots; p.iters = c.iters; p.nwork = c.nwork;
    p.kw_magic = (65536 + g.KW - 1) / g.KW;
    p.v_dt = c.ns / g.Cg; p.v_dq = c.ns % g.Cg;
    p.flip = reverse ? (g.orient ^ 3) : g.orient;
    p.csize = c.csize;
    dim3 grid(c.grid_x, g.groups);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(c.threads);
    cfg.dynamicSmemBytes = c.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = c.csize > 1 ? 1 : 0;
#define IFK_LAUNCH(KERN)                                                                                  \
    {                                                                                                     \
        auto kern = KERN;                                                                                 \
        if (c.smem_bytes > 48 * 1024) {                                                                   \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                                 (int)c.smem_bytes);                                      \
            if (e != cudaSuccess) return (int)e;                                                          \
        }                                                                                                 \
        return cuda_status(cudaLaunchKernelEx(&cfg, kern, p));                                            \
    }
#define X(CC, NV)                                                                                         \
    if (c.cc == CC && c.nv == NV) {                                                                       \
        if (c.csize > 1) IFK_LAUNCH((solve_stream_kernel<CC, NV, true>))                                  \
        else IFK_LAUNCH((solve_stream_kernel<CC, NV, false>))                                             \
    }
    IFK_STREAM_VARIANTS
#undef X
#undef IFK_LAUNCH
    return IFK_ERR_UNSUPPORTED;
}

}  // namespace ifk
