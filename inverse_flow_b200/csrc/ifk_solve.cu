// Wavefront triangular solve  out = L^-1 in  (reverse == false)  or  out = L^-T in
// (reverse == true): host-side variant selection and launch, plus the fallback kernel.
//
//   solve_smem_kernel<CC,NV,VEC>  (ifk_solve_kernel.cuh) image resident in shared memory
//   solve_global_kernel           any shape; neighbours re-read from the output tensor (L1/L2)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <tuple>
#include "ifk_env.cuh"
#include "ifk_solve_kernel.cuh"

namespace ifk {

int solve_variant_max_threads_vec1(int cc, int nv);
int solve_variant_max_threads_vec2(int cc, int nv);
int solve_variant_max_threads_vec4(int cc, int nv);

int solve_variant_max_threads(int cc, int nv, int vec)
{
    switch (vec) {
        case 1: return solve_variant_max_threads_vec1(cc, nv);
        case 2: return solve_variant_max_threads_vec2(cc, nv);
        case 4: return solve_variant_max_threads_vec4(cc, nv);
    }
    return 0;
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
    const int sw = (p.flip & 1) ? -1 : 1, sh = (p.flip & 2) ? -1 : 1;
    const int idx0 = ((p.flip & 2) ? (H - 1) * W : 0) + ((p.flip & 1) ? W - 1 : 0);
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
                const int gr = idx0 + sh * h * W + sw * w;        // memory index of solver pixel (h, w)
                float acc = 0.f;
                for (int ci = 0; ci < Cg; ci++) acc = fmaf(__ldg(wr + ci), in[ci * HW + gr], acc);
                for (int t = 1; t < K; t++) {
                    const int qh = t / KW, qw = t - qh * KW;
                    if (h - qh < 0 || w - qw < 0) continue;
                    const int gn = gr - sh * qh * W - sw * qw;
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
    int cc, nv, vec, ns, nct, nslots, iters, threads, nwork;
    int WP, PS, YN, XN, CgV, NVT, CgP4;
    size_t smem_bytes;
    int grid_x;
};

static int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) l++; return l; }

static bool parse_forced(int *cc, int *nv, int *vec, int *ns, int *nslots)
{
    const EnvKnobs &k = env();                 // IFK_SOLVE_CFG="cc,nv,vec,ns,nslots" -- tuning experiments only
    if (!k.has_solve_cfg) return false;
    *cc = k.solve_cfg[0]; *nv = k.solve_cfg[1]; *vec = k.solve_cfg[2]; *ns = k.solve_cfg[3]; *nslots = k.solve_cfg[4];
    return true;
}

static SolveConfig choose_config_uncached(const Geometry &g)
{
    SolveConfig best{};
    best.smem = false;
    best.threads = 512;
    best.grid_x = g.B < 4 * kNumSM ? g.B : 4 * kNumSM;
    if (best.grid_x < 1) best.grid_x = 1;
    const EnvKnobs &knobs = env();              // testing: pin the fallback / stream / window kernel
    if (knobs.solve_global || knobs.solve_stream || knobs.solve_window) return best;

    const int HP = g.H + g.KH - 1, WP = g.W + g.KW - 1;
    const int XN = round_up(g.Cg * g.H * g.W, 4);
    const int CgP4 = round_up(g.Cg, 4);
    const int nx = g.Cg > 1 ? 2 : 1;
    const int rows = g.H;
    double best_cost = 1e30;
    int fcc, fnv, fvec, fns, fslots;
    const bool forced = parse_forced(&fcc, &fnv, &fvec, &fns, &fslots);
    static const int kVecs[] = {4, 2, 1};
    static const int kNVs[] = {1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 16, 24};
    for (int vec : kVecs) {
        const int CgV = (g.Cg + vec - 1) / vec;
        int PS = CgV * vec;
        if (((PS / vec) & 1) == 0) PS += vec;      // odd stride in vector units: fewer bank conflicts
        const int NVT = (g.K - 1) * CgV;
        const int YN = round_up(HP * WP * PS, 4);
        const size_t smem_bytes = 16 + ((size_t)(g.Cg > 1 ? g.Cg * CgP4 : 0) + (size_t)nx * XN + YN) * sizeof(float);
        if (smem_bytes > (size_t)kMaxSmemBytes) continue;
        const double lane_waste = (double)(CgV * vec) / g.Cg;     // padded channels still cost FMAs
        static const int kCCs[] = {1, 2, 3, 4, 6, 8, 12};
        for (int cc : kCCs) {
            if (cc > g.Cg) continue;
            const int nct = (g.Cg + cc - 1) / cc;
            for (int nv : kNVs) {
                const int tmax = solve_variant_max_threads(cc, nv, vec);
                if (tmax == 0) continue;
                for (int ns = 1; ns <= 32; ns *= 2) {
                    if ((long)ns * nv < NVT) continue;
                    if (ns > 1 && (long)(ns / 2) * nv >= NVT && NVT > 0) continue;     // needless split
                    bool tighter = false;                                              // a smaller nv fits
                    for (int n2 : kNVs)
                        tighter |= (n2 < nv && (long)ns * n2 >= NVT && solve_variant_max_threads(cc, n2, vec) > 0);
                    if (tighter) continue;
                    const int per_slot = ns * nct;
                    if (per_slot > tmax) continue;
                    int nslots = tmax / per_slot;
                    if (nslots > rows) nslots = rows;
                    if (forced) {
                        if (cc != fcc || nv != fnv || vec != fvec || ns != fns) continue;
                        if (fslots > 0 && fslots <= nslots) nslots = fslots;
                    }
                    for (; nslots >= 1; nslots = forced ? 0 : nslots / 2) {
                        const int threads = round_up(nslots * per_slot, 32);
                        const int iters = (rows + nslots - 1) / nslots;
                        const int warps = threads / 32;
                        // per-diagonal cost model (cycles), calibrated with tools/tune_solve.py:
                        // issue slots, shared-memory wavefronts (a 128-bit load costs 4 per warp
                        // whatever it broadcasts) and the dependent latency chain
                        const double live = warps > 1 ? 0.6 * warps : 1.0;       // ~half the rows are on the front
                        const int red = ns > 1 ? 2 * cc : 0;                     // reduce-scatter shuffles (~2*CC)
                        const double instr = nv * (2.0 + cc * vec) + 2.5 * red + 3.0 * cc + 40.0;
                        const double issue = instr * live / 4.0;
                        const double lsu = live * (nv * vec * 1.5 + red + 3.0 * cc);
                        const double chain = (cc >= 4 ? nv * vec * cc : nv * vec * 4.0);
                        const double latency = 80.0 + chain + 35.0 * ilog2(ns) + (warps > 1 ? 30.0 : 0.0);
                        const double waste = (double)(nct * cc) / g.Cg * lane_waste;
                        const double busy = issue > lsu ? issue : lsu;
                        double cost = iters * (busy > latency ? busy + 0.3 * latency : latency + 0.3 * busy);
                        cost *= 1.0 + 0.05 * (waste - 1.0);
                        if (cost < best_cost) {
                            best_cost = cost;
                            best.smem = true;
                            best.cc = cc; best.nv = nv; best.vec = vec; best.ns = ns; best.nct = nct;
                            best.nslots = nslots; best.iters = iters; best.threads = threads;
                            best.WP = WP; best.PS = PS; best.YN = YN; best.XN = XN; best.CgV = CgV;
                            best.NVT = NVT; best.CgP4 = CgP4; best.smem_bytes = smem_bytes;
                        }
                        if (nslots == 1) break;
                    }
                }
            }
        }
    }
    if (!best.smem) return best;
    // helper warps: staging, the pre-pass and the zero fill are latency bound with few threads
    best.nwork = best.threads;
    {
        const int tmax = solve_variant_max_threads(best.cc, best.nv, best.vec);
        int want = 128;
        if (want > tmax) want = tmax;
        if (best.threads < want) best.threads = want;
    }
    // CTAs resident per SM (shared memory and thread limits), then one stripe of images each
    int per_sm = (int)((size_t)(kMaxSmemBytes + 1024) / (best.smem_bytes + 1024));
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

// memoised per geometry: the search above costs 20-30 us on the host, a launch must not
static SolveConfig choose_config(const Geometry &g)
{
    typedef std::tuple<int, int, int, int, int, int, int> Key;
    static std::map<Key, SolveConfig> cache;
    static std::mutex mu;
    static unsigned seen_generation = 0;
    const unsigned gen = env_generation();
    const int bcap = 4 * kNumSM;
    const Key key(g.Cg, g.H, g.W, g.KH, g.KW, g.groups, g.B < bcap ? g.B : bcap);
    std::lock_guard<std::mutex> lock(mu);
    if (seen_generation != gen) {                                     // knobs reloaded (tests): start over
        cache.clear();
        seen_generation = gen;
    }
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const SolveConfig c = choose_config_uncached(g);
    cache[key] = c;
    return c;
}

bool solve_use_pdl()
{
    // programmatic dependent launch of the solves: the next solve's prologue (weights -> registers,
    // zero fill) overlaps the current solve's wavefront.  IFK_PDL=0 switches it off.
    return env().pdl;
}

// which kernel serves an image that is not shared-memory resident: 2 = window (ring of diagonals
// in shared memory), 1 = stream (neighbours through L1/L2; rings that exceed shared memory),
// 0 = the plain fallback.  IFK_SOLVE_GLOBAL / IFK_SOLVE_STREAM = 1 pin the older kernels (tests).
static int large_image_kernel(const Geometry &g)
{
    const EnvKnobs &k = env();
    if (k.solve_global) return 0;
    if (!k.solve_stream && window_solve_available(g)) return 2;
    if (stream_solve_available(g)) return 1;
    return 0;
}

int launch_solve(const Geometry &g, const float *in, const float *prepared, float *out,
                 bool reverse, cudaStream_t s, long long *probe)
{
    if (g.B == 0) return 0;
    const float *prep_dir = prepared_dir(g, prepared, reverse ? 1 : 0);
    if (shfl_solve_available(g)) return launch_solve_shfl(g, in, prep_dir, out, reverse, probe, s);
    if (split_solve_available(g)) {
        const float *pack = prepared + split_pack_offset(g);
        return launch_split_layers(g, 1, nullptr, &pack, in, &out, reverse, g.flags, probe, s);
    }
    if (wave_solve_available(g)) return launch_solve_wave(g, in, prepared, out, reverse, g.flags, probe, s);
    const SolveConfig c = choose_config(g);
    SolveParams p{};
    p.in = in; p.out = out; p.prep = prep_dir;
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W; p.KH = g.KH; p.KW = g.KW;
    p.Cg = g.Cg; p.KD = g.KD; p.KDP = g.KDP;
    p.flip = reverse ? (g.orient ^ 3) : g.orient;     // the adjoint walks the fully reflected frame
    p.probe = probe;
    p.early = (g.flags & IFK_FLAG_STABLE_PREPARED) ? 1 : 0;
    dim3 grid(c.grid_x, g.groups);
    if (!c.smem) {
        switch (large_image_kernel(g)) {
            case 2: return launch_solve_window(g, in, prep_dir, out, reverse, s);
            case 1: return launch_solve_stream(g, in, prep_dir, out, reverse, s);
        }
        solve_global_kernel<<<grid, c.threads, 0, s>>>(p);
        return cuda_status(cudaGetLastError());
    }
    p.WP = c.WP; p.PS = c.PS; p.YN = c.YN; p.XN = c.XN; p.CgV = c.CgV; p.NVT = c.NVT; p.CgP4 = c.CgP4;
    p.NS = c.ns; p.NCT = c.nct; p.nslots = c.nslots; p.iters = c.iters; p.nwork = c.nwork;
    p.kw_magic = (65536 + g.KW - 1) / g.KW;
    p.v_dt = c.ns / c.CgV; p.v_dq = c.ns % c.CgV;
    const size_t img_bytes = (size_t)g.Cg * g.H * g.W * sizeof(float);
    p.bulk = (img_bytes % 16 == 0) && (((uintptr_t)in | (uintptr_t)out) % 16 == 0) ? 1 : 0;
    if (env().nobulk) p.bulk = 0;
    p.walign = ((uintptr_t)prep_dir % 16 == 0) ? 1 : 0;
    switch (c.vec) {
        case 1: return launch_solve_vec1(c.cc, c.nv, p, grid, c.threads, c.smem_bytes, s);
        case 2: return launch_solve_vec2(c.cc, c.nv, p, grid, c.threads, c.smem_bytes, s);
        case 4: return launch_solve_vec4(c.cc, c.nv, p, grid, c.threads, c.smem_bytes, s);
    }
    return IFK_ERR_UNSUPPORTED;
}

int describe_solve(const Geometry &g, char *buf, size_t buflen)
{
    if (shfl_solve_available(g)) return describe_shfl_solve(g, buf, buflen);
    if (split_solve_available(g)) return describe_split_solve(g, buf, buflen);
    if (wave_solve_available(g)) return describe_wave_solve(g, buf, buflen);
    const SolveConfig c = choose_config(g);
    if (c.smem)
        snprintf(buf, buflen, "smem<cc=%d,nv=%d,vec=%d> ns=%d nct=%d slots=%d iters=%d threads=%d(%d) smem=%zuB grid=%dx%d",
                 c.cc, c.nv, c.vec, c.ns, c.nct, c.nslots, c.iters, c.threads, c.nwork, c.smem_bytes, c.grid_x, g.groups);
    else {
        switch (large_image_kernel(g)) {
            case 2: return describe_window_solve(g, buf, buflen);
            case 1: return describe_stream_solve(g, buf, buflen);
        }
        snprintf(buf, buflen, "global threads=%d grid=%dx%d", c.threads, c.grid_x, g.groups);
    }
    return 0;
}

}  // namespace ifk
