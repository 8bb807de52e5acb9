// Wavefront solve for images that do not fit in shared memory: the "window" kernel.
//
// An anti-diagonal wavefront only ever looks back KH+KW-2 diagonals, so what has to be on chip
// is not the image but a RING of the last few diagonals.  The ring lives in shared memory in
// skewed coordinates, ring[slot = d mod S][row h][channel] holding pixel (h, d - h): a neighbour
// (h - qh, w - qw) of the pixel on diagonal d sits at slot (d - qh - qw) mod S, row h - qh -- a
// constant offset per reduction entry apart from the wrap of the slot index.  Rows above the
// image (halo) and positions left of it (row h > diagonal d': never written since the ring was
// zeroed) read as zero, so the walk needs no border tests.
//
// Same thread decomposition as the resident kernel (ifk_solve_kernel.cuh): thread = (row slot,
// tile of CC output channels, slice `ks` of the (K-1)*Cg neighbour reduction), the slice's
// prepared weights in registers for the whole batch stripe, VEC-wide ld.shared of neighbour
// channels, reduce-scatter over warp shuffles, one barrier per diagonal.  Around it:
//   * pre-pass z = T x, pointwise and coalesced, straight into the OUTPUT tensor (as the older
//     stream kernel does);
//   * IO warps (the last `nio` threads) stage z of diagonal d + kWinPF from the output tensor
//     into the ring slot that diagonal will use, with 4-byte cp.async copies that stay in flight
//     across the barriers: global-memory latency never meets the wavefront's critical path;
//   * the lane that finishes a channel adds the staged z, writes y into the ring (for the next
//     diagonals) and fire-and-forget into the output tensor.
// The wavefront therefore reads shared memory only: a warp-wide gather of 32 different pixels
// costs one conflict-free shared-memory wavefront instead of 32 L1 tag look-ups, which is what
// bounded the stream kernel (ifk_solve_stream.cu).
//
// Cluster mode (CL): a thread-block cluster of 2..16 CTAs shares one image when one SM's
// register file cannot hold the weights.  Each CTA owns a slice of the OUTPUT channels; the y
// values it finishes are written into the ring of every CTA of the cluster through distributed
// shared memory (mapa + st.shared::cluster) and the per-diagonal barrier becomes
// barrier.cluster (arrive.release / wait.acquire).
//
// Replaces, for large images, the same reference loop as the resident kernel
// (inv_conv_with_bp_kernel_general.cu:72-129).
#include <stdio.h>
#include <stdlib.h>
#include "ifk_env.cuh"
#include "ifk_solve_kernel.cuh"

namespace ifk {

constexpr int kWinPF = 2;       // diagonals of z in flight ahead of the front

struct WindowParams {
    const float *in;
    float *out;
    const float *prep;
    int B, C, H, W, KH, KW, Cg, KDP, CgP4;
    int CgV;            // channel vectors per pixel = ceil(Cg / VEC)
    int NVT;            // (K-1) * CgV vector entries of one output's reduction
    int PS;             // ring pixel stride in floats
    int HPr;            // ring rows per slot = H + KH - 1 (zero halo on top)
    int S;              // ring slots = KH + KW - 1 + kWinPF
    int NS, NCT, nslots, iters, nwork, nio;
    int kw_magic, v_dt, v_dq;
    int flip;           // as SolveParams::flip
    int csize;          // CTAs per cluster (1 = none); CTA r owns channel tiles [r*NCT, (r+1)*NCT)
};

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void win_cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t win_cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// y into the ring of every CTA of the cluster (own CTA included) through distributed shared memory
__device__ __forceinline__ void sts_cluster_all(uint32_t addr, float v, int csize)
{
    for (int r = 0; r < csize; r++) {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(r));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
    }
}

// reduce-scatter as Rs (ifk_solve_kernel.cuh) for RP independent pixels at once (their shuffles
// interleave); the finishing lane adds z, writes the ring and HBM
template <int N, int LEVELS, bool CL, int RP>
struct RsWin {
    __device__ __forceinline__ static void run(float (*acc)[N], float (*zv)[N], int ks, int m, int own_size,
                                               const bool *active, const uint32_t *ya, float *const *gp, int gstride,
                                               int csize)
    {
        static_assert(LEVELS >= 0, "");
        if (m == 0 || LEVELS == 0) {
#pragma unroll
            for (int r = 0; r < RP; r++)
#pragma unroll
                for (int i = 0; i < N; i++)
                    if (active[r] && i < own_size) {
                        const float yv = acc[r][i] + zv[r][i];
                        if (CL) sts_cluster_all(ya[r] + 4u * i, yv, csize);
                        else sts_f32(ya[r] + 4u * i, yv);
                        gp[r][(size_t)i * gstride] = yv;
                    }
            return;
        }
        constexpr int HALF = (N + 1) / 2;
        const bool hi = (ks & m) != 0;
        float nxt[RP][HALF], nz[RP][HALF];
#pragma unroll
        for (int r = 0; r < RP; r++)
#pragma unroll
            for (int i = 0; i < HALF; i++) {
                const float lo_v = acc[r][i];
                const float hi_v = (i + HALF < N) ? acc[r][i + HALF] : 0.f;
                nxt[r][i] = (hi ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, hi ? lo_v : hi_v, m);
                nz[r][i] = zv[r][i];
            }
        RsWin<HALF, (LEVELS > 0 ? LEVELS - 1 : 0), CL, RP>::run(nxt, nz, ks, m >> 1, own_size, active, ya, gp, gstride,
                                                               csize);
    }
};

template <int CC, int NV, int VEC, int RP>
constexpr int window_max_threads()
{
    // CC*NV*VEC weights + per pixel in flight (NV*VEC loaded values + sums) + NV offsets + packed slot
    // phases + bookkeeping
    int regs = CC * NV * VEC + RP * (NV * VEC + 3 * CC) + NV + (NV + 7) / 8 + 68;
    if (regs > 255) regs = 255;
    int t = (65536 / regs) / 32 * 32;
    return t > 1024 ? 1024 : t;
}

// RP = rows a thread keeps in flight per step: with a single row slot per CTA (wide groups) a step is
// one long dependent chain (loads -> FMAs -> shuffles -> stores); two independent pixels interleave
template <int CC, int NV, int VEC, bool CL, int RP>
__global__ void __launch_bounds__(window_max_threads<CC, NV, VEC, RP>())
solve_window_kernel(const WindowParams p)
{
    extern __shared__ __align__(16) float smem[];
    const int Cg = p.Cg, H = p.H, W = p.W, HW = p.H * p.W, S = p.S, PS = p.PS, HPr = p.HPr;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int G = blockIdx.y;
    const float *wg = p.prep + (size_t)G * Cg * p.KDP;
    float *tT = smem;                                        // [Cg][CgP4] transposed T (pre-pass)
    float *ring = smem + Cg * p.CgP4;                        // [S][HPr][PS]
    const uint32_t ring_u32 = smem_u32(ring);
    const int ring_floats = S * HPr * PS;
    const uint32_t slot_bytes = (uint32_t)(HPr * PS) * 4u, ring_bytes = slot_bytes * (uint32_t)S;
    const uint32_t row_step = (uint32_t)(p.nslots * PS) * 4u;   // next row iteration of a thread

    const int NS = p.NS, NCT = p.NCT, KH1 = p.KH - 1;
    const int crank = CL ? (int)win_cluster_ctarank() : 0;
    const bool is_work = tid < p.nwork;
    const int ks = tid % NS;
    const int ct = crank * NCT + (tid / NS) % NCT;           // global channel tile of this thread
    const int slot = tid / (NS * NCT);
    const bool worker = is_work && slot < p.nslots;
    const int srow = worker ? slot : 0;

    // this thread's slice of the prepared kernel -> registers; ring address of each entry at d = 0
    float wreg[CC][NV * VEC];
    uint32_t offs[NV];
    uint32_t spack[(NV + 7) / 8];        // 4 bits per entry: s = qh + qw, the slot distance
#pragma unroll
    for (int i = 0; i < (NV + 7) / 8; i++) spack[i] = 0u;
    {
        int t1 = ks / p.CgV, q = ks - t1 * p.CgV;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const bool valid = worker && j * NS + ks < p.NVT;
            const int t = valid ? t1 + 1 : 1, qq = valid ? q : 0;
            const int qh = (t * p.kw_magic) >> 16, qw = t - qh * p.KW;
            const int s = qh + qw;
            offs[j] = ring_u32 + (uint32_t)((((S - s) * HPr) + (srow - qh + KH1)) * PS + qq * VEC) * 4u;
            spack[j / 8] |= (uint32_t)s << (4 * (j % 8));
            const int wcol = valid ? t1 * Cg + q * VEC : 0;
#pragma unroll
            for (int e = 0; e < VEC; e++)
#pragma unroll
                for (int cc = 0; cc < CC; cc++) {
                    const int co = ct * CC + cc;
                    wreg[cc][j * VEC + e] = (valid && co < Cg && qq * VEC + e < Cg)
                                                ? __ldg(wg + (size_t)co * p.KDP + Cg + wcol + e) : 0.f;
                }
            q += p.v_dq;
            t1 += p.v_dt;
            if (q >= p.CgV) { q -= p.CgV; t1++; }
        }
    }
    for (int i = tid; i < Cg * p.CgP4; i += nthr) {
        const int ci = i / p.CgP4, co = i - ci * p.CgP4;
        tT[i] = co < Cg ? __ldg(wg + (size_t)co * p.KDP + ci) : 0.f;
    }

    int own_off, own_size;
    rs_owner(CC, NS, ks, &own_off, &own_size);
    {
        const int tile_n = Cg - ct * CC < CC ? Cg - ct * CC : CC;
        own_size = own_off + own_size > tile_n ? (tile_n - own_off > 0 ? tile_n - own_off : 0) : own_size;
    }
    const int own_c0 = ct * CC + own_off;
    const int ndiag = H + W - 1;
    // memory index of solver pixel (h, w) = idx0 + sh_w*h + sw_1*w (reflected axes walk backwards)
    const int sw_1 = (p.flip & 1) ? -1 : 1, sh_w = (p.flip & 2) ? -W : W;
    const int idx0 = ((p.flip & 2) ? (H - 1) * W : 0) + ((p.flip & 1) ? W - 1 : 0);

    // IO threads: the channels whose z this CTA stages (its own output slice) and the walk over
    // (row, channel) items in steps of nio without divisions
    const int c_lo = CL ? crank * NCT * CC : 0;
    int c_n = CL ? NCT * CC : Cg;
    if (c_lo + c_n > Cg) c_n = Cg - c_lo > 0 ? Cg - c_lo : 0;
    const int io_tid = tid - p.nwork;
    const int cdiv = c_n > 0 ? c_n : 1;
    const int io_r0 = io_tid >= 0 ? io_tid / cdiv : 0, io_c0 = io_tid >= 0 ? io_tid - io_r0 * cdiv : 0;
    const int io_dr = p.nio / cdiv, io_dc = p.nio - io_dr * cdiv;

    const int nclusters = CL ? gridDim.x / p.csize : gridDim.x;
    for (int b = CL ? blockIdx.x / p.csize : blockIdx.x; b < p.B; b += nclusters) {
        const size_t gbase = ((size_t)b * p.C + (size_t)G * Cg) * HW;
        const float *in_b = p.in + gbase;
        float *out_b = p.out + gbase;

        // ring back to zero (halo rows, and the "left of the image" positions rely on it)
        for (int i = tid * 4; i < ring_floats; i += nthr * 4)
            *reinterpret_cast<float4 *>(ring + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b == (CL ? blockIdx.x / p.csize : blockIdx.x)) __syncthreads();      // tT complete (first image)

        // pre-pass z = T x, pointwise, coalesced over the pixels; z goes straight to the output
        const int n4 = p.CgP4 >> 2;
        for (int i = (CL ? crank * nthr : 0) + tid; i < HW * n4; i += (CL ? p.csize : 1) * nthr) {
            const int c4 = i / HW, r = i - c4 * HW;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const float *tp = tT + c4 * 4;
#pragma unroll 8
            for (int ci = 0; ci < Cg; ci++) {          // 8 loads in flight per thread: the pre-pass is latency bound
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
        if (CL) win_cluster_barrier();   // z visible cluster-wide; every ring of the cluster is zero
        else __syncthreads();

        // stage z of diagonal dn into ring slot sn (IO threads; one cp.async group per diagonal)
        auto stage = [&](int dn, int sn) {
            if (dn < ndiag && io_tid >= 0) {
                const int hlo = dn - (W - 1) > 0 ? dn - (W - 1) : 0;
                const int hhi = dn < H - 1 ? dn : H - 1;
                const int items = (hhi - hlo + 1) * c_n;
                int r = io_r0, c = io_c0;
                for (int e = io_tid; e < items; e += p.nio) {
                    const int h = hlo + r, w = dn - h;
                    const float *src = out_b + (size_t)(c_lo + c) * HW + (idx0 + sh_w * h + sw_1 * w);
                    cp_async4(ring_u32 + (uint32_t)(((sn * HPr) + h + KH1) * PS + c_lo + c) * 4u, src);
                    c += io_dc;
                    r += io_dr;
                    if (c >= c_n) { c -= c_n; r++; }
                }
            }
            cp_async_commit();
        };
        int sn = 0;                      // slot of the next diagonal to stage
#pragma unroll
        for (int k = 0; k < kWinPF; k++) {
            stage(k, sn);
            sn = sn + 1 == S ? 0 : sn + 1;
        }
        cp_async_wait<kWinPF - 1>();     // diagonal 0 has landed
        if (CL) win_cluster_barrier(); else __syncthreads();

        // own-channel slot of row `slot` on diagonal 0; output pointer of that pixel
        uint32_t ya_d = ring_u32 + (uint32_t)((srow + KH1) * PS + own_c0) * 4u;
        float *gp_d = out_b + (size_t)own_c0 * HW + idx0 + srow * (sh_w - sw_1);
        const int g_row = p.nslots * (sh_w - sw_1);
        int u = 0;                       // d mod S
        for (int d = 0; d < ndiag; d++) {
            if (is_work) {
                uint32_t radd = 0u;
                float *gp = gp_d;
                int col = d - srow;
#pragma unroll 1
                for (int it = 0; it < p.iters; it += RP, radd += RP * row_step, gp += RP * g_row, col -= RP * p.nslots) {
                    bool active[RP];
                    uint32_t ya[RP];
                    float *gpp[RP];
                    bool any = false;
#pragma unroll
                    for (int r = 0; r < RP; r++) {
                        active[r] = worker && it + r < p.iters && srow + (it + r) * p.nslots < H &&
                                    (unsigned)(col - r * p.nslots) < (unsigned)W;
                        any |= active[r];
                    }
                    if (!__any_sync(0xffffffffu, any)) continue;           // warp-uniform
                    float v[RP][NV * VEC];
                    float zv[RP][CC];
#pragma unroll
                    for (int r = 0; r < RP; r++) {
                        const uint32_t ra = active[r] ? radd + r * row_step : 0u;   // idle lanes: a legal row
                        ya[r] = ya_d + ra;
                        gpp[r] = gp + r * g_row;
#pragma unroll
                        for (int j = 0; j < NV; j++) lds_vec<VEC>(v[r] + j * VEC, offs[j] + ra);
#pragma unroll
                        for (int i = 0; i < CC; i++) zv[r][i] = (active[r] && i < own_size) ? lds_f32(ya[r] + 4u * i) : 0.f;
                    }
                    constexpr int NACC = (CC >= 4 || RP > 1) ? 1 : (CC >= 2 ? 2 : 4);  // independent FMA chains
                    float acc[RP][CC];
#pragma unroll
                    for (int r = 0; r < RP; r++) {
                        float part[NACC][CC];
#pragma unroll
                        for (int a = 0; a < NACC; a++)
#pragma unroll
                            for (int cc = 0; cc < CC; cc++) part[a][cc] = 0.f;
#pragma unroll
                        for (int i = 0; i < NV * VEC; i++)
#pragma unroll
                            for (int cc = 0; cc < CC; cc++)
                                part[i % NACC][cc] = fmaf(wreg[cc][i], v[r][i], part[i % NACC][cc]);
#pragma unroll
                        for (int cc = 0; cc < CC; cc++) {
                            acc[r][cc] = part[0][cc];
#pragma unroll
                            for (int a = 1; a < NACC; a++) acc[r][cc] += part[a][cc];
                        }
                    }
                    RsWin<CC, 5, CL, RP>::run(acc, zv, ks, NS >> 1, own_size, active, ya, gpp, HW, p.csize);
                }
            } else {
                stage(d + kWinPF, sn);
                cp_async_wait<kWinPF - 1>();     // diagonal d + 1 has landed (this thread's copies)
            }
            if (CL) win_cluster_barrier(); else __syncthreads();
            // diagonal d + 1: every slot index advances by one and wraps at S
            sn = sn + 1 == S ? 0 : sn + 1;
            u = u + 1 == S ? 0 : u + 1;
            ya_d += slot_bytes;
            if (u == 0) ya_d -= ring_bytes;
            gp_d += sw_1;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                offs[j] += slot_bytes;
                if (((spack[j / 8] >> (4 * (j % 8))) & 15u) == (uint32_t)u) offs[j] -= ring_bytes;
            }
        }
        cp_async_wait<0>();
        // the next image starts from diagonal 0 again: step the entry addresses on (no loads) until
        // the slot phase is back at zero
        while (u != 0) {
            u = u + 1 == S ? 0 : u + 1;
#pragma unroll
            for (int j = 0; j < NV; j++) {
                offs[j] += slot_bytes;
                if (((spack[j / 8] >> (4 * (j % 8))) & 15u) == (uint32_t)u) offs[j] -= ring_bytes;
            }
        }
        // the last barrier of the loop ordered every ring read before the zero fill of the next image
    }
}

// ---- host side -----------------------------------------------------------------------------
struct WindowConfig {
    bool ok;
    int cc, nv, vec, rp, ns, nct, nslots, iters, threads, nwork, nio, grid_x;
    int csize, CgV, NVT, PS, HPr, S;
    size_t smem_bytes;
};

// X(CC, NV, RP)
#define IFK_WINDOW_VARIANTS_V4                                                                       \
    X(12, 3, 1) X(8, 3, 1) X(6, 3, 1) X(4, 3, 1) X(8, 5, 1) X(6, 5, 1) X(4, 5, 1) X(6, 6, 1) X(4, 6, 1)        \
    X(3, 6, 1) X(4, 9, 1) X(3, 9, 1) X(2, 9, 1) X(2, 12, 1) X(1, 12, 1) X(1, 18, 1)                             \
    X(8, 3, 2) X(6, 3, 2) X(4, 3, 2) X(4, 5, 2) X(4, 6, 2) X(3, 6, 2) X(2, 9, 2) X(1, 12, 2)
#define IFK_WINDOW_VARIANTS_V1                                                                       \
    X(1, 8, 1) X(1, 12, 1) X(1, 24, 1) X(3, 6, 1) X(3, 12, 1) X(3, 24, 1) X(2, 6, 1) X(2, 12, 1)

static int window_variant_threads(int cc, int nv, int vec, int rp)
{
#define X(CC, NV, RP) if (vec == 4 && cc == CC && nv == NV && rp == RP) return window_max_threads<CC, NV, 4, RP>();
    IFK_WINDOW_VARIANTS_V4
#undef X
#define X(CC, NV, RP) if (vec == 1 && cc == CC && nv == NV && rp == RP) return window_max_threads<CC, NV, 1, RP>();
    IFK_WINDOW_VARIANTS_V1
#undef X
    return 0;
}

static WindowConfig choose_window(const Geometry &g)
{
    WindowConfig best{};
    best.ok = false;
    if (g.K < 2) return best;                       // 1x1 kernel: no neighbours, nothing to stream
    if (g.KH + g.KW - 2 > 15) return best;          // slot distance is packed in 4 bits
    const int vec = g.Cg >= 4 ? 4 : 1;
    const int CgV = (g.Cg + vec - 1) / vec;
    int PS = CgV * vec;
    if (((PS / vec) & 1) == 0) PS += vec;           // odd stride in vector units: conflict-free gathers
    const int NVT = (g.K - 1) * CgV;
    const int HPr = g.H + g.KH - 1, S = g.KH + g.KW - 1 + kWinPF;
    const size_t smem = ((size_t)g.Cg * round_up(g.Cg, 4) + (size_t)round_up(S * HPr * PS, 4)) * sizeof(float);
    if (smem > (size_t)kMaxSmemBytes) return best;
    double best_cost = 1e30;
    int fcc = 0, fnv = 0, fcs = 0, frp = 0;
    fcc = env().window_cfg[0]; fnv = env().window_cfg[1]; fcs = env().window_cfg[2]; frp = env().window_cfg[3];   // IFK_WINDOW_CFG, tuning only
    static const int kCCs[] = {12, 8, 6, 4, 3, 2, 1};
    static const int kNVs[] = {3, 5, 6, 8, 9, 12, 18, 24};
    static const int kCsizes[] = {1, 2, 4, 8, 16};
    const int ndiag = g.H + g.W - 1;
    for (int csize : kCsizes) {
        if (fcs && csize != fcs) continue;
        if (csize > 1 && vec != 4) continue;
        for (int cc : kCCs) {
            if (cc > g.Cg) continue;
            const int nct_total = (g.Cg + cc - 1) / cc;
            if (csize > nct_total) continue;
            const int nct = (nct_total + csize - 1) / csize;
            for (int nv : kNVs)
                for (int rp = 1; rp <= 2; rp++) {
                    const int tmax = window_variant_threads(cc, nv, vec, rp);
                    if (tmax == 0) continue;
                    if (fcc && (cc != fcc || nv != fnv)) continue;
                    if (frp && rp != frp) continue;
                    for (int ns = 1; ns <= 32; ns *= 2) {
                        if ((long)ns * nv < NVT) continue;
                        if (ns > 1 && (long)(ns / 2) * nv >= NVT) continue;
                        const int per_slot = ns * nct;
                        const int nio = 64;
                        if (per_slot + nio > tmax) continue;
                        int nslots = (tmax - nio) / per_slot;
                        if (nslots > g.H) nslots = g.H;
                        const int iters = (g.H + nslots - 1) / nslots;
                        nslots = (g.H + iters - 1) / iters;              // same passes, fewer idle slots
                        if (rp > 1 && iters < 2) continue;
                        const int nwork = round_up(nslots * per_slot, 32);
                        if (nwork + nio > tmax) continue;
                        // Cycles per image, fitted to tools/tune_window.sh runs.  A step (one row per slot, RP
                        // of them interleaved) takes the longer of its dependent chain -- loads, FMAs, shuffle
                        // levels, stores: ~400 + 19 log2(ns) + 1.9 FMAs per thread -- and of what the SM can
                        // issue for the nslots pixels in flight (instructions / 4 schedulers + ~12 cycles per
                        // warp-wide 128-bit gather).  A cluster adds its remote stores (~10 cycles each).
                        int lg = 0;
                        while ((1 << (lg + 1)) <= ns) lg++;
                        const double fmas = (double)cc * nv * vec;
                        const int red = ns > 1 ? 2 * cc : 0;
                        const double instr = nv * (2.0 + cc * vec) + 2.5 * red + 3.0 * cc + 40.0;
                        const double pixel = per_slot / 32.0 * (instr / 4.0 + nv * (vec == 4 ? 12.0 : 3.0));
                        const double chain = 400.0 + 19.0 * lg + 1.9 * fmas;
                        const double chain_rp = rp == 1 ? chain : 1.6 * chain;    // two pixels: measured 0.8x-0.97x per pixel
                        const double tp = (double)rp * nslots * pixel;
                        const double step = chain_rp > tp ? chain_rp : tp;
                        const double waste = (double)(ns * nv) / NVT * (double)(nct * csize * cc) / g.Cg;
                        const double exchange = csize > 1 ? 400.0 + 5.0 * g.H * g.Cg * (1.0 + 0.1 * csize) : 60.0;
                        const double prepass = (double)g.H * g.W * ((g.Cg + 3) / 4) / (nwork + nio) * ((g.Cg + 7) / 8) * 700.0 /
                                               (csize > 1 ? csize : 1);
                        const double per_image = ndiag * (((iters + rp - 1) / rp) * step * (1.0 + 0.05 * (waste - 1.0)) + exchange) +
                                                 prepass;
                        int per_sm = (int)((size_t)(kMaxSmemBytes + 1024) / (smem + 1024));     // CTAs one SM can hold
                        if (per_sm > tmax / (nwork + nio)) per_sm = tmax / (nwork + nio);        // register file
                        if (per_sm > 4) per_sm = 4;
                        if (per_sm < 1) per_sm = 1;
                        const long images = (long)(g.B > 0 ? g.B : 1) * g.groups;
                        const long ctas = images * csize;
                        const int want = (int)((ctas + kNumSM - 1) / kNumSM);
                        const int share = want < per_sm ? want : per_sm;
                        const long resident = (long)kNumSM * share / csize > 0 ? (long)kNumSM * share / csize : 1;
                        const double waves = (double)((images + resident - 1) / resident);
                        // CTAs sharing an SM share its issue slots: a wave of them takes longer than one alone
                        // (two rows in flight also hide the remote-store latency of a cluster: measured 0.8x)
                        const double cost = per_image * waves * (1.0 + 0.6 * (share - 1)) * (rp > 1 ? (csize > 1 ? 0.85 : 0.95) : 1.0);
                        if (cost < best_cost) {
                            best_cost = cost;
                            best.ok = true;
                            best.cc = cc; best.nv = nv; best.vec = vec; best.rp = rp; best.ns = ns; best.nct = nct;
                            best.nslots = nslots; best.iters = iters; best.nwork = nwork; best.nio = nio;
                            best.threads = nwork + nio; best.csize = csize;
                        }
                    }
                }
        }
    }
    if (!best.ok) return best;
    best.CgV = CgV; best.NVT = NVT; best.PS = PS; best.HPr = HPr; best.S = S;
    best.smem_bytes = smem;
    int per_sm = (int)((size_t)(kMaxSmemBytes + 1024) / (smem + 1024));
    const int tmax_best = window_variant_threads(best.cc, best.nv, best.vec, best.rp);
    if (per_sm > tmax_best / best.threads) per_sm = tmax_best / best.threads;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int nclusters = (kNumSM * per_sm / best.csize + g.groups - 1) / g.groups;
    if (nclusters > g.B) nclusters = g.B;
    if (nclusters < 1) nclusters = 1;
    best.grid_x = nclusters * best.csize;
    return best;
}

bool window_solve_available(const Geometry &g) { return choose_window(g).ok; }

int describe_window_solve(const Geometry &g, char *buf, size_t buflen)
{
    const WindowConfig c = choose_window(g);
    snprintf(buf, buflen,
             "window<cc=%d,nv=%d,vec=%d,rp=%d> cluster=%d ns=%d nct=%d slots=%d iters=%d threads=%d+%d ring=%dx%dx%d smem=%zuB grid=%dx%d",
             c.cc, c.nv, c.vec, c.rp, c.csize, c.ns, c.nct, c.nslots, c.iters, c.nwork, c.nio, c.S, c.HPr, c.PS,
             c.smem_bytes, c.grid_x, g.groups);
    return 0;
}

int launch_solve_window(const Geometry &g, const float *in, const float *prep_dir, float *out,
                        bool reverse, cudaStream_t s)
{
    const WindowConfig c = choose_window(g);
    if (!c.ok) return IFK_ERR_UNSUPPORTED;
    WindowParams p{};
    p.in = in; p.out = out; p.prep = prep_dir;
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W; p.KH = g.KH; p.KW = g.KW; p.Cg = g.Cg; p.KDP = g.KDP;
    p.CgP4 = round_up(g.Cg, 4);
    p.CgV = c.CgV; p.NVT = c.NVT; p.PS = c.PS; p.HPr = c.HPr; p.S = c.S;
    p.NS = c.ns; p.NCT = c.nct; p.nslots = c.nslots; p.iters = c.iters; p.nwork = c.nwork; p.nio = c.nio;
    p.kw_magic = (65536 + g.KW - 1) / g.KW;
    p.v_dt = c.ns / c.CgV; p.v_dq = c.ns % c.CgV;
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
        if (c.csize > 8) {                                                                                \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);\
            if (e != cudaSuccess) return (int)e;                                                          \
        }                                                                                                 \
        return cuda_status(cudaLaunchKernelEx(&cfg, kern, p));                                            \
    }
#define X(CC, NV, RP)                                                                                     \
    if (c.vec == 4 && c.cc == CC && c.nv == NV && c.rp == RP) {                                           \
        if (c.csize > 1) IFK_LAUNCH((solve_window_kernel<CC, NV, 4, true, RP>))                           \
        else IFK_LAUNCH((solve_window_kernel<CC, NV, 4, false, RP>))                                      \
    }
    IFK_WINDOW_VARIANTS_V4
#undef X
#define X(CC, NV, RP)                                                                                     \
    if (c.vec == 1 && c.cc == CC && c.nv == NV && c.rp == RP) IFK_LAUNCH((solve_window_kernel<CC, NV, 1, false, RP>))
    IFK_WINDOW_VARIANTS_V1
#undef X
#undef IFK_LAUNCH
    return IFK_ERR_UNSUPPORTED;
}

}  // namespace ifk
