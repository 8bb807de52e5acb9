// Warp-specialised wavefront solve ("split" kernel) for the reference models' mid-sized layers.
//
// Same job as the wave kernel (ifk_solve_wave.cu) -- one launch instead of the reference's
// (H+W-1)*C/4 launches + cudaDeviceSynchronize (inv_conv_with_bp_kernel_general.cu:72-129; adjoint
// .cu:388-483).  What bounded that kernel (profiles/r02_ncu_wave_100x12x16.txt): one warp per scheduler
// issues, IN ORDER, the dependent chain of a diagonal (barrier -> 2 fresh taps -> two shuffle levels ->
// store) AND the three quarters of the arithmetic that do not depend on the newest diagonal; 365 cycles
// per step for 193 cycles of FMA pipe.  Here the two kinds of work live in different warps, so the
// hardware scheduler -- not the instruction order of one warp -- overlaps them:
//
//  * HELPER warps: lane = (image row, tile of CCH output channels, slice ks of NSH of the reduction).
//    During step d they form, for the row's pixel on diagonal d+1, the OLD part: every tap two or more
//    diagonals back plus T x (tap 0 of the prepared kernel, reading the input image) -- 7/9 of the work at
//    k = 3.  No shuffles: each lane leaves its CCH partial sums in a small shared buffer.
//  * CHAIN warps: lane = (image row, CCC output channels).  During step d they finish the row's pixel on
//    diagonal d: the two FRESH taps (0,1), (1,0) over all input channels (128-bit broadcast loads), plus
//    the helpers' NSH partial sums (128-bit loads), one store.  No shuffle, no cross-lane reduction: the
//    dependent chain of a step is  barrier -> shared loads -> NFT*CG/2 packed FMAs per channel -> store.
//
// One block barrier per diagonal.  Everything else (TMA bulk load of the NCHW image, transposition into
// zero-haloed NHWC buffers, orientation by index reflection, programmatic dependent launch, consecutive
// layers in one launch) is as in the wave kernel.
#include <map>
#include <mutex>
#include <stdio.h>
#include <tuple>
#include "ifk_env.cuh"
#include "ifk_solve_kernel.cuh"

namespace ifk {

constexpr int kSplitChainMax = 8;      // consecutive layers one launch can take (ifk_inverse_chain_f32)

struct SplitParams {
    const float *in;
    float *out[kSplitChainMax];           // layer i's output (every layer's y reaches memory: the backward needs it)
    const float4 *pack[kSplitChainMax];   // layer i's packed weights of this direction: [group][helper | chain]
    int flips[kSplitChainMax];            // layer i's frame (as SolveParams::flip)
    int nlayers;
    int B, C, H, W;
    int bulk, early;
    int PS, RSP;        // pixel stride / row stride of the NHWC buffers, floats (multiples of 4)
    int YN, XN;         // floats per NHWC buffer / of the NCHW staging buffer
    int PRS;            // row stride of the partial-sum buffer, floats
    unsigned mW;        // ceil(2^32 / W)
    int tm_shift;       // log2 of the pixel lanes of the transposing passes
    int dbg;            // tuning aid (IFK_SPLIT_CFG second field): bit 0 skips the helpers' work, bit 1 the chain's
    long long *probe;
};

template <int CG, int KH, int KW, int CCH, int NSH, int CCC, int NROW>
struct SplitCfg {
    static_assert(CG % CCH == 0 && CG % CCC == 0 && CG % 4 == 0, "tiles must divide the group");
    static_assert(NSH % 4 == 0, "the partial sums of a channel are read as 128-bit words");
    static_assert(KH > 1 && KW > 1, "two fresh taps");
    static constexpr int K = KH * KW;
    static constexpr int NCTH = CG / CCH;
    static constexpr int LPPH = NCTH * NSH;                // helper lanes per pixel
    static constexpr int CGV = CG / 2;                     // packed pairs per tap
    static constexpr int NFT = 2;                          // fresh taps (0,1), (1,0)
    static constexpr int NOT = K - 1 - NFT;                // older y taps
    static constexpr int NO = (NOT + 1) * CGV;             // old pair entries: y taps + the T (input) tap
    static constexpr int NVO = (NO + NSH - 1) / NSH;       // per helper lane
    static constexpr int NW4H = (NVO * CCH + 1) / 2;       // float4 of packed weights per helper lane
    static constexpr int LPC = CG / CCC;                   // chain lanes per pixel
    static constexpr int NPC = NFT * CCC * CGV;            // packed pairs per chain lane
    static constexpr int NW4C = (NPC + 1) / 2;
    static constexpr int NHT = NROW * LPPH;                // helper threads
    static_assert(NHT % 32 == 0, "roles are warp-uniform");
    static constexpr int NCT = (NROW * LPC + 31) / 32 * 32;   // chain threads
    static constexpr int NTHR = NHT + NCT;
    static constexpr int NWREG = 2 * (NW4H > NW4C ? NW4H : NW4C);
    static constexpr int GS4 = NW4H * LPPH + NW4C * LPC;   // float4 per (direction, group)
    static constexpr int PART_ROW = CG * NSH;              // floats of one row's partial sums (before padding)
};

#define IFK_SPROBE(i) do { if (p.probe && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.probe[i] = clock64(); } while (0)

__device__ __forceinline__ f32x2_t add_f32x2(f32x2_t a, f32x2_t b)
{
    f32x2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// the ti-th tap (1 <= t < K) that lies two or more diagonals back
template <int KH, int KW>
__host__ __device__ inline int split_old_tap(int ti)
{
    int n = 0;
#pragma unroll
    for (int tt = 1; tt < KH * KW; tt++) {
        if (tt / KW + tt % KW < 2) continue;
        if (n == ti) return tt;
        n++;
    }
    return 0;
}

template <int CG, int KH, int KW, int CCH, int NSH, int CCC, int NROW>
__global__ void __launch_bounds__((SplitCfg<CG, KH, KW, CCH, NSH, CCC, NROW>::NTHR))
solve_split_kernel(const SplitParams p)
{
    typedef SplitCfg<CG, KH, KW, CCH, NSH, CCC, NROW> Cfg;
    constexpr int LPPH = Cfg::LPPH, CGV = Cfg::CGV, NVO = Cfg::NVO, NW4H = Cfg::NW4H, NW4C = Cfg::NW4C;
    constexpr int LPC = Cfg::LPC, NHT = Cfg::NHT, nthr = Cfg::NTHR, NOT = Cfg::NOT, NO = Cfg::NO;
    constexpr int VEC = 2;                              // channel vector of the transposing passes
    IFK_SPROBE(0);
    extern __shared__ __align__(128) float smem[];
    const int H = p.H, W = p.W, HW = p.H * p.W, PS = p.PS, RSP = p.RSP;
    const int tid = threadIdx.x;

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);         // 16 bytes reserved
    float *xbuf = smem + 4;                                     // [CG][HW] the input image as it lies in memory
    float *yb = xbuf + p.XN;                                    // [H+KH-1][..][PS] y, zero halo top / left
    float *xh = yb + p.YN;                                      // same geometry: the input image, NHWC
    float *part = xh + p.YN;                                    // [2][NROW][PRS] the helpers' partial sums

    const int G = blockIdx.y;
    const uint32_t img_bytes = (uint32_t)(CG * HW) * 4u;
    const size_t img_stride = (size_t)p.C * HW;
    const float *in0 = p.in + (size_t)G * CG * HW;

    const bool is_helper = tid < NHT;                           // warp-uniform
    // helper: (row slot, channel tile ct, reduction slice ks); chain: (row slot, channel lane lc)
    const int l = tid % LPPH, ks = l % NSH, ct = l / NSH;
    const int ctid = tid - NHT, lc = is_helper ? 0 : ctid % LPC;
    const int slot = is_helper ? tid / LPPH : ctid / LPC;
    const bool worker = slot < H && (is_helper || ctid < NROW * LPC);

    f32x2_t wreg[Cfg::NWREG];
    int offs[NVO];
    const int xoff = p.YN * 4;                                  // xh lies YN floats behind yb
    auto load_weights = [&](int li) {
        const ulonglong2 *pk = reinterpret_cast<const ulonglong2 *>(p.pack[li]) + (size_t)G * Cfg::GS4;
        if (is_helper) {
            pk += l;
#pragma unroll
            for (int j = 0; j < NW4H; j++) {
                const ulonglong2 w4 = __ldg(pk + j * LPPH);
                wreg[2 * j] = w4.x;
                wreg[2 * j + 1] = w4.y;
            }
        } else {
            pk += NW4H * LPPH + lc;
#pragma unroll
            for (int j = 0; j < NW4C; j++) {
                const ulonglong2 w4 = __ldg(pk + j * LPC);
                wreg[2 * j] = w4.x;
                wreg[2 * j + 1] = w4.y;
            }
        }
    };
    // byte offset (from the pixel's own position in yb) of the neighbour pair each old entry of this lane reads
#pragma unroll
    for (int j = 0; j < NVO; j++) {
        const int e = j * NSH + ks;
        int off = 0;
        if (e < NO) {
            const int ti = e / CGV, q = e - ti * CGV;
            if (ti == NOT) off = q * 8 + xoff;                  // the T tap reads the input image
            else {
                const int t = split_old_tap<KH, KW>(ti);
                off = (-(t / KW) * RSP - (t % KW) * PS) * 4 + q * 8;
            }
        }
        offs[j] = off;                                          // padding entries: offset 0, zero weights
    }

    // Programmatic dependent launch (see the wave kernel): the weights may be fetched ahead of the
    // dependency wait only when the caller vouches nobody is writing them (IFK_FLAG_STABLE_PREPARED).
    if (p.early) load_weights(0);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int b = blockIdx.x;
    if (p.bulk && tid == 0) {
        mbar_init(bar, 1);
        if (b < p.B) {
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b * img_stride, img_bytes, bar);
        }
    }
    if (tid == 0) smem[2] = 0.f;          // source of the opaque zero used by Hold
    if (!p.early) load_weights(0);
    // zero halo: once per CTA -- the interior of xh is rewritten for every image, the interior of yb is
    // written before it is read (padding entries read it with zero weights: it must stay finite)
    for (int i = tid * 4; i < 2 * p.YN; i += nthr * 4)
        *reinterpret_cast<float4 *>(yb + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();            // zero fill and mbarrier init visible
    IFK_SPROBE(1);

    Hold hold;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hold.zero) : "r"(smem_u32(smem + 2)) : "memory");
    const uint32_t ybase = hold(smem_u32(yb) + (uint32_t)(((KH - 1) * RSP + (KW - 1) * PS) * 4));   // pixel (0, 0)
    const uint32_t pix_step = hold((uint32_t)PS * 4u);                                   // per diagonal
    const uint32_t pix0 = ybase + (uint32_t)(slot * (RSP - PS)) * 4u;                    // row `slot`, d = 0
    const int ndiag = hold(H + W - 1);
    const int Wr = hold(W);
    const int slot_r = hold(slot);
    // partial sums of row `slot`: [parity][row][cc][ks / 4][lc][ks % 4]
    const uint32_t part0 = smem_u32(part) + (uint32_t)(slot * p.PRS) * 4u;
    const uint32_t part_par = hold((uint32_t)(NROW * p.PRS) * 4u);
    // helper: where partial `cc` of this lane goes (channel c = ct*CCH + cc -> chain lane c / CCC, its channel c % CCC)
    uint32_t pst[CCH];
#pragma unroll
    for (int cc = 0; cc < CCH; cc++) {
        const int c = ct * CCH + cc, clane = c / CCC, ccc = c % CCC;
        pst[cc] = part0 + (uint32_t)(((ccc * (NSH / 4) + ks / 4) * LPC + clane) * 4 + (ks % 4)) * 4u;
    }
    const uint32_t pld = part0 + (uint32_t)(lc * 4) * 4u;       // chain: first 128-bit word of its partial sums
    const uint32_t fr_w = hold((uint32_t)(PS * 4)), fr_h = hold((uint32_t)(RSP * 4));
    const int TM = 1 << p.tm_shift, TC = nthr >> p.tm_shift;
    const int tm = tid & (TM - 1), tc = tid >> p.tm_shift;
    IFK_SPROBE(2);

    uint32_t parity = 0;
    for (; b < p.B; b += gridDim.x) {
        const int b_next = b + gridDim.x;
        IFK_SPROBE(3);
        if (p.bulk) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            const float *src = in0 + (size_t)b * img_stride;
            for (int i = tid; i < CG * HW; i += nthr) xbuf[i] = __ldg(src + i);
            __syncthreads();
        }
        IFK_SPROBE(4);
        // transpose the image into xh (pixel lanes x channel lanes, as in the wave kernel)
        for (int m = tm; m < HW; m += TM) {
            const int hm = (int)__umulhi((unsigned)m, p.mW), wm = m - hm * W;
            const int h = (p.flips[0] & 2) ? H - 1 - hm : hm, w = (p.flips[0] & 1) ? W - 1 - wm : wm;
            float *d = xh + ((h + KH - 1) * RSP + (w + KW - 1) * PS) + tc * VEC;
            const float *sp = xbuf + m + tc * VEC * HW;
            const int dstep = TC * VEC, sstep = TC * VEC * HW;
#pragma unroll 2
            for (int cv = tc; cv < CGV; cv += TC, d += dstep, sp += sstep)
                *reinterpret_cast<float2 *>(d) = make_float2(sp[0], sp[HW]);
        }
        __syncthreads();
        if (p.bulk && tid == 0 && b_next < p.B) {                 // xbuf is free: prefetch the next image
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b_next * img_stride, img_bytes, bar);
        }
        IFK_SPROBE(5);

      for (int li = 0; li < p.nlayers; li++) {      // consecutive layers: the image stays in shared memory
        // ---- wavefront ------------------------------------------------------------------------
        // helper: old part (taps two or more diagonals back + T x) of the row's pixel on diagonal d1
        auto helper_step = [&](int d1) {
            const int col = d1 - slot_r;
            const bool act = worker && (unsigned)col < (unsigned)Wr;
            if (!__any_sync(0xffffffffu, act)) return;             // warp-uniform
            const uint32_t pn = act ? pix0 + (uint32_t)d1 * pix_step : ybase;
            f32x2_t v[NVO];
#pragma unroll
            for (int j = 0; j < NVO; j++) lds_pairs<2>(v + j, pn + (uint32_t)offs[j]);
            f32x2_t a[CCH];
#pragma unroll
            for (int cc = 0; cc < CCH; cc++) a[cc] = 0ull;
#pragma unroll
            for (int i = 0; i < NVO; i++)
#pragma unroll
                for (int cc = 0; cc < CCH; cc++) a[cc] = fma_f32x2(wreg[i * CCH + cc], v[i], a[cc]);
            const uint32_t po = (d1 & 1) ? part_par : 0u;
#pragma unroll
            for (int cc = 0; cc < CCH; cc++) sts_f32(pst[cc] + po, sum_f32x2(a[cc]));   // (idle rows: never read)
        };
        // chain: the row's pixel on diagonal d = fresh taps over all input channels + the helpers' partial sums
        auto chain_step = [&](int d) {
            const int col = d - slot_r;
            const bool act = worker && (unsigned)col < (unsigned)Wr;
            if (!__any_sync(0xffffffffu, act)) return;             // warp-uniform
            const uint32_t pa = act ? pix0 + (uint32_t)d * pix_step : ybase;
            f32x2_t vf[2][CGV];
#pragma unroll
            for (int q = 0; q < CG / 4; q++) lds_pairs<4>(&vf[0][2 * q], pa - fr_w + 16u * q);      // tap (0,1)
#pragma unroll
            for (int q = 0; q < CG / 4; q++) lds_pairs<4>(&vf[1][2 * q], pa - fr_h + 16u * q);      // tap (1,0)
            const uint32_t pl = pld + ((d & 1) ? part_par : 0u);
            float4 pr[CCC][NSH / 4];
#pragma unroll
            for (int cc = 0; cc < CCC; cc++)
#pragma unroll
                for (int k4 = 0; k4 < NSH / 4; k4++) pr[cc][k4] = lds_f32x4(pl + (uint32_t)((cc * (NSH / 4) + k4) * LPC * 16));
            f32x2_t acc[CCC][2];
#pragma unroll
            for (int cc = 0; cc < CCC; cc++) acc[cc][0] = acc[cc][1] = 0ull;
#pragma unroll
            for (int i = 0; i < CGV; i++)
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int cc = 0; cc < CCC; cc++)
                        acc[cc][t] = fma_f32x2(wreg[(t * CCC + cc) * CGV + i], vf[t][i], acc[cc][t]);
            float yv[CCC];
#pragma unroll
            for (int cc = 0; cc < CCC; cc++) {
                float ps = 0.f;
#pragma unroll
                for (int k4 = 0; k4 < NSH / 4; k4++) ps += (pr[cc][k4].x + pr[cc][k4].y) + (pr[cc][k4].z + pr[cc][k4].w);
                yv[cc] = sum_f32x2(add_f32x2(acc[cc][0], acc[cc][1])) + ps;
            }
            if (act) {
                const uint32_t ya = pa + (uint32_t)(lc * CCC) * 4u;
                if (CCC == 2) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ya), "f"(yv[0]), "f"(yv[CCC - 1]) : "memory");
                else {
#pragma unroll
                    for (int cc = 0; cc < CCC; cc++) sts_f32(ya + 4u * cc, yv[cc]);
                }
            }
        };

        const bool do_h = is_helper && !(p.dbg & 1), do_c = !is_helper && !(p.dbg & 2);
        if (do_h) helper_step(0);
        __syncthreads();
        for (int d = 0; d < ndiag; d++) {
            if (is_helper) {
                if (do_h && d + 1 < ndiag) helper_step(d + 1);
            } else if (do_c) {
                chain_step(d);
            }
            __syncthreads();                    // diagonal d and the old parts of diagonal d+1 are visible
        }
        IFK_SPROBE(6);

        // ---- y leaves: NHWC shared memory -> NCHW global, coalesced.  In a chain the same pass hands y to the
        //      next layer: into xh, re-indexed from this layer's frame to the next one's.
        const bool more = li + 1 < p.nlayers;
        if (more || (p.nlayers > 1 && b_next < p.B)) load_weights(more ? li + 1 : 0);      // in flight during the write-out
        float *dst = p.out[li] + (size_t)G * CG * HW + (size_t)b * img_stride;
        const int fl = p.flips[li], fn = more ? p.flips[li + 1] : fl;
        for (int m = tm; m < HW; m += TM) {
            const int hm = (int)__umulhi((unsigned)m, p.mW), wm = m - hm * W;
            const int h = (fl & 2) ? H - 1 - hm : hm, w = (fl & 1) ? W - 1 - wm : wm;
            const int h2 = (fn & 2) ? H - 1 - hm : hm, w2 = (fn & 1) ? W - 1 - wm : wm;
            const float *sp = yb + ((h + KH - 1) * RSP + (w + KW - 1) * PS) + tc * VEC;
            float *xn = xh + ((h2 + KH - 1) * RSP + (w2 + KW - 1) * PS) + tc * VEC;
            float *d = dst + m + tc * VEC * HW;
            const int sstep = TC * VEC, dstep = TC * VEC * HW;
#pragma unroll 2
            for (int cv = tc; cv < CGV; cv += TC, sp += sstep, xn += sstep, d += dstep) {
                const float2 t2 = *reinterpret_cast<const float2 *>(sp);
                d[0] = t2.x; d[HW] = t2.y;
                if (more) *reinterpret_cast<float2 *>(xn) = t2;
            }
        }
        if (more) __syncthreads();                 // xh holds the next layer's input; yb may be overwritten
      }
        IFK_SPROBE(7);
        if (b_next < p.B) __syncthreads();          // yb is rewritten by the next image's wavefront
    }
    IFK_SPROBE(8);
}

// Packed weights of the split kernel, built from the canonical prepared rows ([co][tap][ci], ifk_prepare.cu).
// Per (layer, direction, group): helper section [NW4H][LPPH] float4 -- flat pair f = i*CCH + cc of lane
// (ct, ks) is entry e = i*NSH + ks (tap, channel pair) of output channel ct*CCH + cc -- then the chain
// section [NW4C][LPC] float4 -- flat pair f = (t*CCC + cc)*CGV + i: fresh tap t, channel lc*CCC + cc, pair i.
struct SplitPackParams {
    const float *prepared;   // canonical, [layer][dir][group][co][KDP]
    float *pack;             // [layer][dir][group][GS4] float4
    size_t prepared_stride, pack_stride;      // floats between layers
    int C, cg, kh, kw, cch, nsh, ccc, KDP, groups, count;
};

__global__ void __launch_bounds__(256)
split_pack_kernel(const SplitPackParams q)
{
    const int K = q.kh * q.kw, cgv = q.cg / 2, ncth = q.cg / q.cch, lpph = ncth * q.nsh, lpc = q.cg / q.ccc;
    const int not_ = K - 3, NO = (not_ + 1) * cgv, nvo = (NO + q.nsh - 1) / q.nsh;
    const int nw4h = (nvo * q.cch + 1) / 2, npc = 2 * q.ccc * cgv, nw4c = (npc + 1) / 2;
    const int gs4 = nw4h * lpph + nw4c * lpc;
    const long per_group = (long)gs4 * 2;                       // pairs
    const long total = (long)q.count * 2 * q.groups * per_group;
    for (long e0 = blockIdx.x * (long)blockDim.x + threadIdx.x; e0 < total; e0 += (long)gridDim.x * blockDim.x) {
        long r = e0;
        const int pi = (int)(r % per_group); r /= per_group;     // pair index inside the group's block
        const int G = (int)(r % q.groups); r /= q.groups;
        const int dir = (int)(r % 2);
        const int layer = (int)(r / 2);
        const int w4 = pi / 2, half = pi & 1;                    // float4 number, which pair of it
        float w0 = 0.f, w1 = 0.f;
        const float *rows = q.prepared + (size_t)layer * q.prepared_stride + ((size_t)dir * q.C + (size_t)G * q.cg) * q.KDP;
        if (w4 < nw4h * lpph) {                                  // helper section
            const int j4 = w4 / lpph, l = w4 - j4 * lpph;
            const int f = 2 * j4 + half;                         // flat pair of the lane
            const int ks = l % q.nsh, ct = l / q.nsh;
            if (f < nvo * q.cch) {
                const int i = f / q.cch, cc = f - i * q.cch;
                const int ent = i * q.nsh + ks;
                if (ent < NO) {
                    const int ti = ent / cgv, qv = ent - ti * cgv;
                    int t = 0;                                   // ti == not_: the T tap
                    if (ti < not_) {
                        int n = 0;
                        for (int tt = 1; tt < K; tt++) {
                            if (tt / q.kw + tt % q.kw < 2) continue;
                            if (n == ti) { t = tt; break; }
                            n++;
                        }
                    }
                    const float *src = rows + (size_t)(ct * q.cch + cc) * q.KDP + t * q.cg + qv * 2;
                    w0 = src[0];
                    w1 = src[1];
                }
            }
        } else {                                                 // chain section
            const int w4c = w4 - nw4h * lpph;
            const int j4 = w4c / lpc, lc = w4c - j4 * lpc;
            const int f = 2 * j4 + half;
            if (f < npc) {
                const int i = f % cgv, tc = f / cgv;
                const int cc = tc % q.ccc, tf = tc / q.ccc;
                const int t = tf == 0 ? 1 : q.kw;                // (0,1), (1,0)
                const float *src = rows + (size_t)(lc * q.ccc + cc) * q.KDP + t * q.cg + i * 2;
                w0 = src[0];
                w1 = src[1];
            }
        }
        float *dstp = q.pack + (size_t)layer * q.pack_stride + (((size_t)dir * q.groups + G) * gs4 + w4) * 4 + half * 2;
        dstp[0] = w0;
        dstp[1] = w1;
    }
}

// ---- host side -----------------------------------------------------------------------------
// X(CG, KH, KW, CCH, NSH, CCC, NROW)
#define IFK_SPLIT_VARIANTS                                  \
    X(12, 3, 3, 6, 4, 2, 16) X(12, 3, 3, 6, 8, 2, 16)      \
    X(24, 3, 3, 6, 8, 2, 8)

struct SplitVariant {
    int cg, kh, kw, cch, nsh, ccc, nrow;
};
static const SplitVariant kSplitVariants[] = {
#define X(CG, KHc, KWc, CCH, NSH, CCC, NROW) {CG, KHc, KWc, CCH, NSH, CCC, NROW},
    IFK_SPLIT_VARIANTS
#undef X
};

struct SplitDims {
    int lpph, lpc, nw4h, nw4c, gs4, nht, nthr;
    size_t pack_floats;      // per layer: both directions, all groups
};
static SplitDims split_dims(const SplitVariant &v, int groups)
{
    SplitDims d{};
    const int K = v.kh * v.kw, cgv = v.cg / 2;
    const int NO = (K - 3 + 1) * cgv, nvo = (NO + v.nsh - 1) / v.nsh;
    d.lpph = (v.cg / v.cch) * v.nsh;
    d.lpc = v.cg / v.ccc;
    d.nw4h = (nvo * v.cch + 1) / 2;
    d.nw4c = (2 * v.ccc * cgv + 1) / 2;
    d.gs4 = d.nw4h * d.lpph + d.nw4c * d.lpc;
    d.nht = v.nrow * d.lpph;
    d.nthr = d.nht + round_up(v.nrow * d.lpc, 32);
    d.pack_floats = (size_t)2 * groups * d.gs4 * 4;
    return d;
}

// the variant family serving a (Cg, KH, KW): fixes the packed-weight layout, so it must not depend on the
// image size or the batch.  IFK_SPLIT_CFG="nsh" picks among the compiled reduction splits (tuning).
static const SplitVariant *split_family(const Geometry &g)
{
    const EnvKnobs &k = env();
    if (k.split_off || k.wave_off || k.pins_other_solver()) return nullptr;
    for (const SplitVariant &v : kSplitVariants) {
        if (v.cg != g.Cg || v.kh != g.KH || v.kw != g.KW) continue;
        if (k.split_cfg[0] && k.split_cfg[0] != v.nsh) continue;
        return &v;
    }
    return nullptr;
}

struct SplitConfig {
    bool ok;
    SplitVariant v;
    int PS, RSP, YN, XN, PRS;
    size_t smem_bytes;
};

static SplitConfig choose_split(const Geometry &g)
{
    SplitConfig c{};
    c.ok = false;
    const SplitVariant *v = split_family(g);
    if (!v || g.H > v->nrow || g.W < 2 || g.H * g.W < 2) return c;
    const SplitDims d = split_dims(*v, g.groups);
    // NHWC layout: pixel stride a multiple of 4 floats (128-bit fresh-tap loads) with PS/4 odd; the row pad
    // keeps the rows of one warp's gathers apart: (RSP - PS) mod 32 in [8, 24]
    int PS = round_up(g.Cg, 4);
    if (((PS / 4) & 1) == 0) PS += 4;
    int RSP = (g.W + g.KW) * PS;
    while (((RSP - PS) % 32 + 32) % 32 < 8 || ((RSP - PS) % 32 + 32) % 32 > 24) RSP += 4;
    // partial sums: row stride with (PRS / 4) mod 8 == LPC mod 8 (consecutive rows continue the bank walk)
    int PRS = g.Cg * v->nsh;
    while (((PRS / 4) % 8) != (d.lpc % 8)) PRS += 4;
    c.v = *v; c.PS = PS; c.RSP = RSP; c.PRS = PRS;
    c.YN = round_up((g.H + g.KH - 1) * RSP, 4);
    c.XN = round_up(g.Cg * g.H * g.W, 4);
    c.smem_bytes = 16 + ((size_t)c.XN + 2 * (size_t)c.YN + (size_t)2 * v->nrow * PRS) * sizeof(float);
    if (c.smem_bytes > (size_t)device_max_smem_optin()) return c;
    c.ok = true;
    return c;
}

bool split_solve_available(const Geometry &g) { return choose_split(g).ok; }

int describe_split_solve(const Geometry &g, char *buf, size_t buflen)
{
    const SplitConfig c = choose_split(g);
    const SplitDims d = split_dims(c.v, g.groups);
    int gx = device_sm_count() / (g.groups > 0 ? g.groups : 1);
    if (gx > g.B) gx = g.B;
    snprintf(buf, buflen, "split<cg=%d,k=%dx%d,helper cc=%d ns=%d,chain cc=%d,rows=%d> threads=%d(%d helper) ps=%d rsp=%d prs=%d "
             "smem=%zuB grid=%dx%d", c.v.cg, c.v.kh, c.v.kw, c.v.cch, c.v.nsh, c.v.ccc, c.v.nrow, d.nthr, d.nht, c.PS, c.RSP,
             c.PRS, c.smem_bytes, gx < 1 ? 1 : gx, g.groups);
    return 0;
}

size_t split_pack_floats(const Geometry &g)
{
    const SplitVariant *v = split_family(g);
    return v ? split_dims(*v, g.groups).pack_floats : 0;
}

// `pack`: the split section of layer 0's prepared buffer
int launch_split_pack(const Geometry &g, const float *prepared, float *pack, int count, size_t prepared_stride, cudaStream_t s)
{
    const SplitVariant *v = split_family(g);
    if (!v || count <= 0) return 0;
    const SplitDims d = split_dims(*v, g.groups);
    SplitPackParams q{};
    q.prepared = prepared;
    q.pack = pack;
    q.prepared_stride = prepared_stride;
    q.pack_stride = prepared_stride;
    q.C = g.C; q.cg = v->cg; q.kh = v->kh; q.kw = v->kw; q.cch = v->cch; q.nsh = v->nsh; q.ccc = v->ccc;
    q.KDP = g.KDP; q.groups = g.groups; q.count = count;
    const long total = (long)count * 2 * g.groups * d.gs4 * 2;
    long blocks = (total + 255) / 256;
    if (blocks > 8L * device_sm_count()) blocks = 8L * device_sm_count();
    split_pack_kernel<<<(unsigned)blocks, 256, 0, s>>>(q);
    return cuda_status(cudaGetLastError());
}

// one launch over `n` consecutive layers (n == 1: a plain solve).  packs[i]: the split section of layer i's prepared buffer
int launch_split_layers(const Geometry &g, int n, const int *orients, const float *const *packs, const float *in,
                        float *const *outs, bool reverse, int flags, long long *probe, cudaStream_t s)
{
    const SplitConfig c = choose_split(g);
    if (!c.ok || n < 1 || n > kSplitChainMax) return IFK_ERR_UNSUPPORTED;
    const SplitDims d = split_dims(c.v, g.groups);
    SplitParams p{};
    p.in = in;
    p.nlayers = n;
    for (int i = 0; i < n; i++) {
        if ((uintptr_t)packs[i] % 16 != 0) return IFK_ERR_UNSUPPORTED;        // read as 16-byte words
        p.pack[i] = reinterpret_cast<const float4 *>(packs[i] + (reverse ? d.pack_floats / 2 : 0));
        p.out[i] = outs[i];
        const int orient = orients ? orients[i] : g.orient;
        p.flips[i] = reverse ? (orient ^ 3) : orient;      // the adjoint walks the fully reflected frame
    }
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W;
    const size_t img_bytes = (size_t)g.Cg * g.H * g.W * sizeof(float);
    p.bulk = (img_bytes % 16 == 0) && ((uintptr_t)in % 16 == 0) && !env().nobulk ? 1 : 0;
    p.early = (flags & IFK_FLAG_STABLE_PREPARED) ? 1 : 0;
    p.PS = c.PS; p.RSP = c.RSP; p.YN = c.YN; p.XN = c.XN; p.PRS = c.PRS;
    p.mW = (unsigned)((0x100000000ULL + (unsigned)g.W - 1) / (unsigned)g.W);
    {   // pixel lanes: the largest power of two <= H*W that divides the thread count
        int sh = 0;
        while ((2 << sh) <= g.H * g.W && d.nthr % (2 << sh) == 0) sh++;
        p.tm_shift = sh;
    }
    p.probe = probe;
    p.dbg = env().split_cfg[1];
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(d.nthr);
    cfg.dynamicSmemBytes = c.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env().pdl ? 1 : 0;
#define X(CG, KHc, KWc, CCH, NSH, CCC, NROW)                                                          \
    if (c.v.cg == CG && c.v.kh == KHc && c.v.kw == KWc && c.v.cch == CCH && c.v.nsh == NSH && c.v.ccc == CCC && \
        c.v.nrow == NROW) {                                                                           \
        auto kern = solve_split_kernel<CG, KHc, KWc, CCH, NSH, CCC, NROW>;                            \
        if (c.smem_bytes > 48 * 1024) {                                                               \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 (int)c.smem_bytes);                                  \
            if (e != cudaSuccess) return (int)e;                                                      \
        }                                                                                             \
        int occ = 1;                                                                                  \
        if (g.B * g.groups > device_sm_count() &&                                                     \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, d.nthr, c.smem_bytes) != cudaSuccess) \
            occ = 1;                                                                                  \
        if (occ < 1) occ = 1;                                                                         \
        int gx = (device_sm_count() * occ + g.groups - 1) / g.groups;                                 \
        if (gx > g.B) gx = g.B;                                                                       \
        cfg.gridDim = dim3(gx < 1 ? 1 : gx, g.groups);                                                \
        return cuda_status(cudaLaunchKernelEx(&cfg, kern, p));                                        \
    }
    IFK_SPLIT_VARIANTS
#undef X
    return IFK_ERR_UNSUPPORTED;
}

}  // namespace ifk
