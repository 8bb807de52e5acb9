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
    int HP, WP, CS;     // halo-padded rows / cols, channel stride (floats) in shared memory
    int NS, NCT, nslots;
    int reverse;
};

// ------------------------------------------------------------------------------------------
// Shared-memory resident kernel.
// thread -> (slot, ct, ks): `slot` = position on the current anti-diagonal, `ct` = tile of CC
// output channels, `ks` = slice of the K*Cg reduction (kidx = j*NS + ks, j < CHUNK).
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
    extern __shared__ __align__(16) float smem[];
    const int Cg = p.Cg, CS = p.CS, WP = p.WP, H = p.H, W = p.W, HW = p.H * p.W;
    float *ybuf = smem;                   // [Cg][CS], zero halo on top/left
    const int XOFF = Cg * CS;             // x buffer sits right behind, same geometry

    const int tid = threadIdx.x;
    const int NS = p.NS, NCT = p.NCT;
    const int ks = tid % NS;
    const int ct = (tid / NS) % NCT;
    const int slot = tid / (NS * NCT);
    const bool worker = slot < p.nslots;
    const int G = blockIdx.y;
    const unsigned lane = tid & 31u;
    const unsigned gmask = NS >= 32 ? 0xffffffffu : (((1u << NS) - 1u) << (lane & ~(unsigned)(NS - 1)));

    // this thread's slice of the prepared kernel -> registers, for the whole batch stripe
    float wreg[CC][CHUNK];
    int offs[CHUNK];
    {
        const float *wg = p.prep + (size_t)G * Cg * p.KDP;
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {
            const int kidx = j * NS + ks;
            const bool valid = worker && kidx < p.KD;
            const int t = valid ? kidx / Cg : 0;
            const int ci = valid ? kidx - t * Cg : 0;
            const int qh = t / p.KW, qw = t - qh * p.KW;
            // padding entries (weight 0) read the pixel's own x value: always finite data of this image
            offs[j] = (!valid || t == 0) ? XOFF + ci * CS : ci * CS - qh * WP - qw;
#pragma unroll
            for (int cc = 0; cc < CC; cc++) {
                const int co = ct * CC + cc;
                wreg[cc][j] = (valid && co < Cg) ? __ldg(wg + (size_t)co * p.KDP + kidx) : 0.f;
            }
        }
    }

    for (int i = tid; i < XOFF; i += blockDim.x) ybuf[i] = 0.f;   // halo stays zero throughout

    const int ndiag = H + W - 1;
    const int halo = (p.KH - 1) * WP + (p.KW - 1);
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const size_t gbase = ((size_t)b * p.C + (size_t)G * Cg) * HW;
        // stage the group's image (contiguous in NCHW): coalesced, reflected if reverse
        for (int i = tid; i < Cg * HW; i += blockDim.x) {
            const int ci = i / HW, r = i - ci * HW;
            const int rr = p.reverse ? HW - 1 - r : r;
            const int h = rr / W, w = rr - h * W;
            ybuf[XOFF + ci * CS + h * WP + w + halo] = __ldg(p.in + gbase + i);
        }
        __syncthreads();

        for (int d = 0; d < ndiag; d++) {
            const int hmin = d - (W - 1) > 0 ? d - (W - 1) : 0;
            const int hmax = d < H - 1 ? d : H - 1;
            if (worker) {
                for (int h = hmin + slot; h <= hmax; h += p.nslots) {
                    const float *px = ybuf + h * WP + (d - h) + halo;
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
                        float *py = ybuf + h * WP + (d - h) + halo + ct * CC * CS;
#pragma unroll
                        for (int cc = 0; cc < CC; cc++)
                            if (ct * CC + cc < Cg) py[cc * CS] = acc0[cc];
                    }
                }
            }
            if (blockDim.x <= 32) __syncwarp(); else __syncthreads();
        }

        for (int i = tid; i < Cg * HW; i += blockDim.x) {
            const int ci = i / HW, r = i - ci * HW;
            const int rr = p.reverse ? HW - 1 - r : r;
            const int h = rr / W, w = rr - h * W;
            p.out[gbase + i] = ybuf[ci * CS + h * WP + w + halo];
        }
        // the next stripe's staging only writes the x buffer, whose readers are all past the
        // last diagonal's barrier; its first diagonal is fenced by the barrier after staging.
    }
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
    int HP, WP, CS;
    size_t smem_bytes;
    int grid_x;
};

static const int kCCs[] = {1, 2, 3, 4};
static const int kChunks[] = {4, 6, 8, 9, 12, 16, 18, 24, 27, 32};

template <int CC, int CHUNK>
static int max_threads_of() { return solve_max_threads<CC, CHUNK>(); }

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
    const size_t smem_bytes = (size_t)2 * g.Cg * CS * sizeof(float);
    const char *force_global = getenv("IFK_SOLVE_GLOBAL");
    if (smem_bytes > (size_t)kMaxSmemBytes || (force_global && force_global[0] == '1')) return best;

    const int diag = g.H < g.W ? g.H : g.W;
    double best_cost = 1e30;
    int fcc, fchunk, fns, fslots;
    const bool forced = parse_forced(&fcc, &fchunk, &fns, &fslots);
    for (int cc : kCCs) {
        if (cc > g.Cg) continue;
        const int nct = (g.Cg + cc - 1) / cc;
        for (int chunk : kChunks) {
            for (int ns = 1; ns <= 32; ns *= 2) {
                if ((long)ns * chunk < g.KD) continue;
                if (ns > 1 && (long)(ns / 2) * chunk >= g.KD) continue;     // needless split
                const int per_slot = ns * nct;
                const int tmax = max_threads_for(cc, chunk);
                if (per_slot > tmax) continue;
                int nslots = tmax / per_slot;
                if (nslots > diag) nslots = diag;
                if (forced) {
                    if (cc != fcc || chunk != fchunk || ns != fns) continue;
                    if (fslots > 0 && fslots <= nslots) nslots = fslots;
                }
                for (; nslots >= 1; nslots = forced ? 0 : nslots / 2) {
                    const int threads = round_up(nslots * per_slot, 32);
                    const int iters = (diag + nslots - 1) / nslots;
                    const int warps = threads / 32;
                    const double instr = chunk * (1.0 + cc) + 2.0 * cc * ilog2(ns) + 24.0;
                    const double waste = (double)(ns * chunk) / g.KD * (double)(nct * cc) / g.Cg;
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
    best.HP = HP; best.WP = WP; best.CS = CS; best.smem_bytes = smem_bytes;
    // CTAs resident per SM (shared memory and thread limits), then one stripe of images each
    int per_sm = (int)((size_t)kMaxSmemBytes / (smem_bytes + 1024));
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
        case 4:  return launch_smem_variant<CC, 4>(p, c, groups, s);
        case 6:  return launch_smem_variant<CC, 6>(p, c, groups, s);
        case 8:  return launch_smem_variant<CC, 8>(p, c, groups, s);
        case 9:  return launch_smem_variant<CC, 9>(p, c, groups, s);
        case 12: return launch_smem_variant<CC, 12>(p, c, groups, s);
        case 16: return launch_smem_variant<CC, 16>(p, c, groups, s);
        case 18: return launch_smem_variant<CC, 18>(p, c, groups, s);
        case 24: return launch_smem_variant<CC, 24>(p, c, groups, s);
        case 27: return launch_smem_variant<CC, 27>(p, c, groups, s);
        case 32: return launch_smem_variant<CC, 32>(p, c, groups, s);
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
    p.HP = c.HP; p.WP = c.WP; p.CS = c.CS; p.NS = c.ns; p.NCT = c.nct; p.nslots = c.nslots;
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
