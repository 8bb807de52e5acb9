// Software-pipelined wavefront solve for the reference models' mid-sized layers ("wave" kernel).
//
// Same job as the resident kernel (ifk_solve_kernel.cuh) -- one launch instead of the reference's
// (H+W-1)*C/4 launches + cudaDeviceSynchronize (inv_conv_with_bp_kernel_general.cu:72-129; adjoint
// .cu:388-483) -- re-organised around what bounded that kernel on the B200 (profiles/r01_*):
//
//  * Fresh / old split.  Of the K-1 neighbour taps of pixel (h, w) on anti-diagonal d only (0,1)
//    and (1,0) lie on diagonal d-1; all others are at least two diagonals old.  A thread therefore
//    computes, inside step d, the OLD part of its row's NEXT pixel (diagonal d+1: everything it
//    reads was visible at the barrier that opened step d) and carries it in registers; the
//    dependent chain of a step shrinks to  barrier -> 2 shared loads -> a quarter of the FMAs ->
//    shuffle reduce -> store,  and three quarters of the arithmetic fills its latency gaps.
//  * T folded.  z = T x (T = (I + A0)^-1, tap 0 of the prepared kernel) is one more "old" tap that
//    reads the input image instead of y: no pre-pass, no z buffer.  The image arrives NCHW by one
//    TMA bulk copy, is transposed once into an NHWC buffer of the same geometry as y (one barrier),
//    and the finished y leaves by coalesced stores straight from the NHWC buffer.
//  * Everything compile-time (group width, taps, channel tile, reduction split, vector width, rows
//    per thread): no index arithmetic, no constant-bank reads, no padding work in the loop.
//  * NHWC buffers with an odd pixel stride (in vectors) and a row pad chosen on the host by
//    enumerating the bank conflicts of the gather pattern (r01: 44 % conflict wavefronts).
//
// Thread -> (row slot, channel tile ct of CC outputs, reduction slice ks of NS); lanes of one pixel
// = NCT*NS.  After the reduce-scatter over the NS lanes each finished channel sits on one lane,
// which writes it.
#include <map>
#include <mutex>
#include <stdio.h>
#include <tuple>
#include "ifk_env.cuh"
#include "ifk_solve_kernel.cuh"

namespace ifk {

constexpr int kWaveChainMax = 8;      // consecutive layers one launch can take (ifk_inverse_chain_f32)

struct WaveParams {
    const float *in;
    float *out[kWaveChainMax];          // layer i's output (every layer's y reaches memory: the backward needs it)
    const float4 *pack[kWaveChainMax];  // layer i's packed weights of this direction (wave_pack_kernel): [group][j][lane]
    int flips[kWaveChainMax];           // layer i's frame (as SolveParams::flip)
    int nlayers;                        // 1 for a plain solve
    const int *codes;    // [slot][lane of the pixel]: tap / channel-vector code per entry (the same for every layer)
    int B, C, H, W;
    int nslots;         // image rows in flight (threads / lanes per pixel)
    int bulk;           // TMA bulk copy of the input image possible (size / alignment)
    int early;          // prepared weights may be fetched ahead of griddepcontrol.wait (ifk.h: IFK_FLAG_*)
    int PS, RSP;        // pixel stride / row stride of the NHWC buffers, floats
    int YN;             // floats per NHWC buffer
    int XN;             // floats of the NCHW staging buffer
    unsigned mW;        // ceil(2^32 / W): m / W == umulhi(m, mW) for every pixel index
    int tm_shift;       // log2 of the pixel lanes of the transposing passes (a power of two dividing the threads)
    long long *probe;   // clock64() stamps of CTA (0,0) thread 0 (ifk_inverse_probe_f32), or nullptr
    // fused neighbours (ifk.h: ifk_fused; single-layer launches only)
    const float *in_scale, *in_bias;   // per-channel affine applied while the image is transposed into xh
    const float *out_scale;            // per-channel scale of the second output
    float *out2;                       // second output (scaled, un-squeezed), or nullptr
    int squeeze_in, squeeze_out;       // the input / the second output is a (C/4, 2H, 2W) image
    int fused;                         // any of the above
};

// reduce-scatter of N per-lane partial sums over M adjacent lanes (compile-time recursive halving,
// see Rs in ifk_solve_kernel.cuh): N/2 + N/4 + ... shuffles, log2(M) levels
template <int N, int M>
struct RsC {
    __device__ __forceinline__ static void run(float *acc, int ks)
    {
        if constexpr (M > 1) {
            constexpr int HALF = (N + 1) / 2;
            const bool hi = (ks & (M / 2)) != 0;
#pragma unroll
            for (int i = 0; i < HALF; i++) {
                const float lo_v = acc[i];
                const float hi_v = (i + HALF < N) ? acc[i + HALF] : 0.f;
                const float send = hi ? lo_v : hi_v;
                const float keep = hi ? hi_v : lo_v;
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, M / 2);
            }
            RsC<HALF, M / 2>::run(acc, ks);
        }
    }
};
__host__ __device__ constexpr int rs_final(int n, int m) { return m > 1 ? rs_final((n + 1) / 2, m / 2) : n; }

// predicated shared-memory store without a branch (a branch around the store splits the step's basic
// block and with it ptxas' freedom to fill the shuffle latencies with the look-ahead FMAs)
__device__ __forceinline__ void sts_f32_if(uint32_t addr, float v, bool pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.f32 [%0], %1;\n\t}"
                 ::"r"(addr), "f"(v), "r"((unsigned)pred) : "memory");
}

template <int CG, int KH, int KW, int CC, int NS, int VEC>
struct WaveCfg {
    static_assert(VEC == 2 || VEC == 4, "packed pairs need an even vector width");
    static_assert(CG % VEC == 0 && CG % CC == 0, "tiles must divide the group");
    static constexpr int K = KH * KW;
    static constexpr int NCT = CG / CC;
    static constexpr int LPP = NCT * NS;                       // lanes per pixel
    static_assert((32 % NS) == 0, "the reduction split stays inside a warp");
    static_assert((LPP % 32) == 0 || (32 % LPP) == 0, "pixels must not straddle warps unevenly");
    static constexpr int CGV = CG / VEC;
    static constexpr int NFT = (KW > 1 ? 1 : 0) + (KH > 1 ? 1 : 0);   // fresh taps: (0,1), (1,0)
    static constexpr int NOT = K - 1 - NFT;                           // older y taps
    static constexpr int NF = NFT * CGV;                               // fresh vector entries
    static constexpr int NO = (NOT + 1) * CGV;                         // old entries: y taps + the T (input) tap
    static constexpr int NVF = (NF + NS - 1) / NS;
    static constexpr int NVO = (NO + NS - 1) / NS;
    static constexpr int PF = NVF * VEC / 2, PO = NVO * VEC / 2;       // packed pairs per output channel
    static constexpr int NW4 = (CC * (PF + PO) + 1) / 2;               // float4 of packed weights per thread
    static constexpr int OWN = rs_final(CC, NS);                       // finished channels per lane (max)
};

// entry code (wave_pack_kernel writes them, the solve decodes them): which neighbour a vector entry reads
__host__ __device__ inline int wave_code(int qh, int qw, int chan, bool is_x) { return qh | (qw << 8) | (chan << 16) | (is_x ? 1 << 30 : 0); }

#define IFK_WPROBE(i) do { if (p.probe && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.probe[i] = clock64(); } while (0)

template <int CG, int KH, int KW, int CC, int NS, int VEC, int ITERS, int NTHR>
__global__ void __launch_bounds__(NTHR)
solve_wave_kernel(const WaveParams p)
{
    typedef WaveCfg<CG, KH, KW, CC, NS, VEC> Cfg;
    constexpr int LPP = Cfg::LPP, CGV = Cfg::CGV, NVF = Cfg::NVF, NVO = Cfg::NVO, PF = Cfg::PF, PO = Cfg::PO;
    constexpr int OWN = Cfg::OWN, NW4 = Cfg::NW4;
    constexpr int TU = VEC == 2 ? 3 : 2;   // vectors per batch of the transposing passes (larger batches cost the loop registers: ptxas then adds moves to it)
    IFK_WPROBE(0);
    extern __shared__ __align__(128) float smem[];
    const int H = p.H, W = p.W, HW = p.H * p.W, PS = p.PS, RSP = p.RSP;
    const int tid = threadIdx.x, nthr = blockDim.x;             // nthr <= NTHR (small images need fewer rows)

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);         // 16 bytes reserved
    float *xbuf = smem + 4;                                     // [CG][HW] the input image as it lies in memory
    float *yb = xbuf + p.XN;                                    // [H+KH-1][..][PS] y, zero halo top / left
    float *xh = yb + p.YN;                                      // same geometry: the input image, NHWC

    const int G = blockIdx.y;
    const uint32_t img_bytes = (uint32_t)(CG * HW) * 4u;
    const size_t img_stride = (size_t)p.C * HW;
    const float *in0 = p.in + (size_t)G * CG * HW;

    const int l = tid % LPP, slot = tid / LPP;
    const int ks = l % NS, ct = l / NS;
    const bool worker = slot < p.nslots;

    // this thread's slice of the prepared kernel -> registers: packed pairs along the reduction axis, laid
    // out by wave_pack_kernel so that a warp reads consecutive 16-byte words (flat pair index i*CC + cc)
    f32x2_t wreg[2 * NW4];
    int offs[NVF + NVO];
    const int xoff = p.YN * 4;                                  // xh lies YN floats behind yb
    auto load_weights = [&](int li) {
        const ulonglong2 *pk = reinterpret_cast<const ulonglong2 *>(p.pack[li]) + (size_t)G * NW4 * LPP + l;
#pragma unroll
        for (int j = 0; j < NW4; j++) {
            const ulonglong2 w4 = __ldg(pk + j * LPP);      // two packed pairs; nothing waits for them here
            wreg[2 * j] = w4.x;
            wreg[2 * j + 1] = w4.y;
        }
    };
    auto load_offsets = [&]() {
#pragma unroll
        for (int j = 0; j < NVF + NVO; j++) {
            const int code = __ldg(p.codes + j * LPP + l);
            const int qh = code & 0xff, qw = (code >> 8) & 0xff, chan = (code >> 16) & 0x3fff;
            offs[j] = ((-qh * RSP - qw * PS) + chan) * 4 + ((code >> 30) & 1) * xoff;
        }
    };

    // Programmatic dependent launch.  Everything that touches only this CTA's shared memory or
    // registers -- mbarrier init, the zero halo, the pinned loop constants -- runs AHEAD of the
    // dependency wait, i.e. while the predecessor on the stream is still inside its wavefront; behind
    // the wait only the image load is left on the launch-to-launch critical path.  The prepared weights
    // may be fetched ahead of the wait only when the caller vouches that the previous operation of the
    // stream did not write them (IFK_FLAG_STABLE_PREPARED); the dependents are released AFTER the wait,
    // so that "the kernel before my predecessor has completed and is visible" holds transitively for
    // them.  Either way the weight fetch (one L2 round trip) is in flight while the image lands and is
    // transposed: nothing below needs the weights before the wavefront starts.
    if (p.early) { load_weights(0); load_offsets(); }
    if (tid == 0) {
        if (p.bulk) mbar_init(bar, 1);
        smem[2] = 0.f;                    // source of the opaque zero used by Hold
    }
    // zero halo (and the extra column the look-ahead touches): once per CTA -- the interior of xh is
    // rewritten for every image, the interior of yb is written before it is read
    for (int i = tid * 4; i < 2 * p.YN; i += nthr * 4)
        *reinterpret_cast<float4 *>(yb + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();            // zero fill and mbarrier init visible

    // which of the tile's CC output channels this lane finishes after the reduce-scatter
    int own_off, own_size;
    rs_owner(CC, NS, ks, &own_off, &own_size);
    if (!worker) own_size = 0;
    Hold hold;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hold.zero) : "r"(smem_u32(smem + 2)) : "memory");
    const uint32_t ybase = hold(smem_u32(yb) + (uint32_t)(((KH - 1) * RSP + (KW - 1) * PS) * 4));   // pixel (0, 0)
    const uint32_t own_bytes = hold((uint32_t)(ct * CC + own_off) * 4u);
    const uint32_t pix_step = hold((uint32_t)PS * 4u);                                   // per diagonal
    const uint32_t pix_iter = hold((uint32_t)(p.nslots * (RSP - PS)) * 4u);              // per row iteration
    const uint32_t pix0 = ybase + (uint32_t)(slot * (RSP - PS)) * 4u;                    // row `slot`, d = 0
    const int ndiag = hold(H + W - 1);
    const int Wr = hold(W), nsl = hold(p.nslots);
    own_size = hold(own_size);
    const int slot_r = hold(slot);
    // memory index of solver pixel (h, w) of a layer's frame = idx0 + sh*h*W + sw*w (reflected axes walk backwards)
    // the transposing passes: TM pixel lanes x TC channel lanes
    const int TM = 1 << p.tm_shift, TC = nthr >> p.tm_shift;
    const int tm = tid & (TM - 1), tc = tid >> p.tm_shift;
    IFK_WPROBE(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    IFK_WPROBE(2);
    int b = blockIdx.x;
    if (p.bulk && tid == 0 && b < p.B) {
        mbar_expect_tx(bar, img_bytes);
        bulk_load(xbuf, in0 + (size_t)b * img_stride, img_bytes, bar);
    }
    if (!p.early) { load_weights(0); load_offsets(); }

    uint32_t parity = 0;
    for (; b < p.B; b += gridDim.x) {
        const int b_next = b + gridDim.x;
        IFK_WPROBE(3);
        if (p.bulk) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            const float *src = in0 + (size_t)b * img_stride;
            for (int i = tid; i < CG * HW; i += nthr) xbuf[i] = __ldg(src + i);
            __syncthreads();
        }
        IFK_WPROBE(4);
        // transpose the image into xh.  Threads = TM pixel lanes x TC channel lanes (TM a power of two): a
        // thread walks memory pixels m = tm, tm + TM, ... and, per pixel, channel vectors cv = tc, tc + TC, ...;
        // consecutive lanes take consecutive pixels, so the loads are conflict free and -- with PS/VEC odd --
        // so are the vector stores.  Loads of up to four vectors are issued before the first store.
        for (int m = tm; m < HW; m += TM) {
            const int hm = (int)__umulhi((unsigned)m, p.mW), wm = m - hm * W;
            const int h = (p.flips[0] & 2) ? H - 1 - hm : hm, w = (p.flips[0] & 1) ? W - 1 - wm : wm;
            float *d = xh + ((h + KH - 1) * RSP + (w + KW - 1) * PS) + tc * VEC;
            const float *sp = xbuf + m + tc * VEC * HW;
            const int dstep = TC * VEC, sstep = TC * VEC * HW;
            if (VEC == 2 && p.fused) {
                // the ActNorm affine and the Squeeze re-indexing ride on this pass: squeezed channel c = 4 c0 + 2 i + j
                // of pixel (hm, wm) is element (c0, 2 hm + i, 2 wm + j) of the (C/4, 2H, 2W) image -- a channel PAIR is
                // two adjacent floats of it
                const float *sc = p.in_scale ? p.in_scale + G * CG : nullptr;
                const float *bi = p.in_bias ? p.in_bias + G * CG : nullptr;
                for (int cv = tc; cv < CGV; cv += TC, d += dstep, sp += sstep) {
                    const int c = cv * 2;
                    float v0, v1;
                    if (p.squeeze_in) {
                        const float2 t2 = *reinterpret_cast<const float2 *>(
                            xbuf + (c >> 2) * 4 * HW + (2 * hm + ((c >> 1) & 1)) * 2 * W + 2 * wm);
                        v0 = t2.x; v1 = t2.y;
                    } else {
                        v0 = sp[0]; v1 = sp[HW];
                    }
                    if (sc) { v0 *= __ldg(sc + c); v1 *= __ldg(sc + c + 1); }
                    if (bi) { v0 += __ldg(bi + c); v1 += __ldg(bi + c + 1); }
                    *reinterpret_cast<float2 *>(d) = make_float2(v0, v1);
                }
                continue;
            }
            // (loads of a whole batch of vectors are issued before its first store: load/store pairs would
            //  serialise on the shared-memory latency, ptxas does not move loads above stores it cannot disambiguate)
            for (int cv = tc; cv < CGV; cv += TU * TC, d += TU * dstep, sp += TU * sstep) {
                float t[TU][VEC];
#pragma unroll
                for (int u = 0; u < TU; u++)
                    if (cv + u * TC < CGV) {
#pragma unroll
                        for (int e = 0; e < VEC; e++) t[u][e] = sp[u * sstep + e * HW];
                    }
#pragma unroll
                for (int u = 0; u < TU; u++)
                    if (cv + u * TC < CGV) {
                        if (VEC == 4) *reinterpret_cast<float4 *>(d + u * dstep) = make_float4(t[u][0], t[u][1], t[u][2], t[u][3]);
                        else *reinterpret_cast<float2 *>(d + u * dstep) = make_float2(t[u][0], t[u][1]);
                    }
            }
        }
        __syncthreads();
        if (p.bulk && tid == 0 && b_next < p.B) {                 // xbuf is free: prefetch the next image
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b_next * img_stride, img_bytes, bar);
        }
        IFK_WPROBE(5);

      for (int li = 0; li < p.nlayers; li++) {      // consecutive layers: the image stays in shared memory
        // ---- wavefront ------------------------------------------------------------------------
        f32x2_t acc[ITERS][CC];          // old part of the row's pixel on the coming diagonal
        // old part (taps two or more diagonals back + T x) of the pixel at `pn`
        auto old_part = [&](uint32_t pn, bool act_n, f32x2_t *a) {
            const uint32_t pa = act_n ? pn : ybase;
            f32x2_t v[PO];
#pragma unroll
            for (int j = 0; j < NVO; j++) lds_pairs<VEC>(v + j * (VEC / 2), pa + (uint32_t)offs[NVF + j]);
#pragma unroll
            for (int cc = 0; cc < CC; cc++) a[cc] = 0ull;
#pragma unroll
            for (int i = 0; i < PO; i++)
#pragma unroll
                for (int cc = 0; cc < CC; cc++) a[cc] = fma_f32x2(wreg[(PF + i) * CC + cc], v[i], a[cc]);
        };
        // one wavefront step of a row: the fresh part of its pixel on this diagonal (taps on the previous
        // diagonal), reduction over the NS lanes, store -- and the old part of its next pixel, whose loads
        // are issued up front and whose FMAs fill the latency of the reduction's shuffles
        auto step = [&](uint32_t pix, bool act, bool act_n, f32x2_t *a) {
            const uint32_t pa = act ? pix : ybase;
            const uint32_t pn = act_n ? pix + pix_step : ybase;
            f32x2_t vf[PF > 0 ? PF : 1], vo[PO];
#pragma unroll
            for (int j = 0; j < NVF; j++) lds_pairs<VEC>(vf + j * (VEC / 2), pa + (uint32_t)offs[j]);
#pragma unroll
            for (int j = 0; j < NVO; j++) lds_pairs<VEC>(vo + j * (VEC / 2), pn + (uint32_t)offs[NVF + j]);
#pragma unroll
            for (int i = 0; i < PF; i++)
#pragma unroll
                for (int cc = 0; cc < CC; cc++) a[cc] = fma_f32x2(wreg[i * CC + cc], vf[i], a[cc]);
            float s[CC];
#pragma unroll
            for (int cc = 0; cc < CC; cc++) s[cc] = sum_f32x2(a[cc]);
#pragma unroll
            for (int cc = 0; cc < CC; cc++) a[cc] = 0ull;
            // (consumed last-loaded-first: every accumulator's chain starts with the last vector, so ptxas has
            //  to put all the loads in flight before the first look-ahead FMA instead of load/use pairs)
#pragma unroll
            for (int i = PO - 1; i >= 0; i--)
#pragma unroll
                for (int cc = 0; cc < CC; cc++) a[cc] = fma_f32x2(wreg[(PF + i) * CC + cc], vo[i], a[cc]);
            RsC<CC, NS>::run(s, ks);
#pragma unroll
            for (int i = 0; i < OWN; i++) sts_f32_if(pa + own_bytes + 4u * i, s[i], act && i < own_size);
        };

        {   // diagonal 0's old part (only pixel (0,0) is live there; the others are discarded)
            uint32_t pix = pix0;
            int col = -slot_r;
#pragma unroll
            for (int it = 0; it < ITERS; it++, pix += pix_iter, col -= nsl) {
                const bool row_ok = worker && slot_r + it * nsl < H;
                old_part(pix, row_ok && (unsigned)col < (unsigned)Wr, acc[it]);
            }
        }
        uint32_t pix_d = pix0;
        for (int d = 0; d < ndiag; d++, pix_d += pix_step) {
            if (d > 0) __syncthreads();             // diagonal d-1 is visible
            uint32_t pix = pix_d;
            int col = d - slot_r;
#pragma unroll
            for (int it = 0; it < ITERS; it++, pix += pix_iter, col -= nsl) {
                const bool row_ok = worker && slot_r + it * nsl < H;
                const bool act = row_ok && (unsigned)col < (unsigned)Wr;
                const bool act_n = row_ok && (unsigned)(col + 1) < (unsigned)Wr;
                if (!__any_sync(0xffffffffu, act || act_n)) continue;       // warp-uniform
                step(pix, act, act_n, acc[it]);
            }
        }
        IFK_WPROBE(6);
        __syncthreads();

        // ---- y leaves: NHWC shared memory -> NCHW global, coalesced (consecutive threads = consecutive
        //      memory pixels of one channel vector; PS/VEC odd keeps the vector loads conflict free).  In a chain the
        //      same pass hands y to the next layer: into xh, re-indexed from this layer's frame to the next one's.
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
            if (VEC == 2 && p.fused) {
                // the adjoint of the fused neighbours: raw result to `out` (dW reads it), per-channel scaled and
                // un-squeezed (depth_to_space) to `out2`
                const float *os = p.out_scale ? p.out_scale + G * CG : nullptr;
                float *dst2 = p.out2 ? p.out2 + (size_t)G * CG * HW + (size_t)b * img_stride : nullptr;
                for (int cv = tc; cv < CGV; cv += TC, sp += sstep, d += dstep) {
                    const float2 t2 = *reinterpret_cast<const float2 *>(sp);
                    if (p.out[li]) { d[0] = t2.x; d[HW] = t2.y; }
                    if (dst2) {
                        const int c = cv * 2;
                        float v0 = t2.x, v1 = t2.y;
                        if (os) { v0 *= __ldg(os + c); v1 *= __ldg(os + c + 1); }
                        if (p.squeeze_out)
                            *reinterpret_cast<float2 *>(dst2 + (c >> 2) * 4 * HW + (2 * hm + ((c >> 1) & 1)) * 2 * W + 2 * wm) =
                                make_float2(v0, v1);
                        else { dst2[c * HW + m] = v0; dst2[(c + 1) * HW + m] = v1; }
                    }
                }
                continue;
            }
            for (int cv = tc; cv < CGV; cv += TU * TC, sp += TU * sstep, xn += TU * sstep, d += TU * dstep) {
                float t[TU][VEC];
#pragma unroll
                for (int u = 0; u < TU; u++)
                    if (cv + u * TC < CGV) {
                        if (VEC == 4) {
                            const float4 t4 = *reinterpret_cast<const float4 *>(sp + u * sstep);
                            t[u][0] = t4.x; t[u][1] = t4.y; t[u][2] = t4.z; t[u][3] = t4.w;
                        } else {
                            const float2 t2 = *reinterpret_cast<const float2 *>(sp + u * sstep);
                            t[u][0] = t2.x; t[u][1] = t2.y;
                        }
                    }
#pragma unroll
                for (int u = 0; u < TU; u++)
                    if (cv + u * TC < CGV) {
#pragma unroll
                        for (int e = 0; e < VEC; e++) d[u * dstep + e * HW] = t[u][e];
                        if (more) {
                            if (VEC == 4) *reinterpret_cast<float4 *>(xn + u * sstep) = make_float4(t[u][0], t[u][1], t[u][2], t[u][3]);
                            else *reinterpret_cast<float2 *>(xn + u * sstep) = make_float2(t[u][0], t[u][1]);
                        }
                    }
            }
        }
        if (more) __syncthreads();                 // xh holds the next layer's input; yb may be overwritten
      }
        IFK_WPROBE(7);
        if (b_next < p.B) __syncthreads();          // yb is rewritten by the next image's wavefront
    }
    IFK_WPROBE(8);
}

// Packed weights for the wave kernel, built from the canonical prepared rows ([co][tap][ci], ifk_prepare.cu):
// for lane l = (ct, ks) of a pixel the flat pair index f = i*CC + cc (i-th packed pair of the lane's reduction
// slice, output channel ct*CC + cc) lands in float4 number f/2 -- stored [f/2][l], so that a warp's loads are
// consecutive 16-byte words.  `codes` gets, once per direction-independent geometry, the neighbour each vector
// entry of a lane reads.  One thread per (layer, dir, group) x (float4, lane): consecutive threads write consecutive
// 16-byte words; the tap of an "old" entry comes from a table the host fills (this kernel sits on the critical path
// in front of a step's first solve: the first version, one thread per pair with a tap search and ~20 integer
// divisions each, took 27 us for 48 layers of Cg = 48).
struct WavePackParams {
    const float *prepared;   // canonical, [layer][dir][group][co][KDP]
    float *pack;             // [layer][dir][group][NW4][LPP] float4
    size_t pack_floats;      // floats of one layer's pack; its codes ([(NVF+NVO)][LPP] ints) follow
    size_t prepared_stride, pack_stride;      // floats between layers (pack_stride: of the pack section)
    int C, cg, kh, kw, cc, ns, vec, KDP, groups, count;
    int aligned;                              // every layer's pack starts at a multiple of 16 bytes
    unsigned char old_tap[64];                // i-th tap (array order) that lies two or more diagonals back
};

__global__ void __launch_bounds__(256)
wave_pack_kernel(const WavePackParams q)
{
    const int K = q.kh * q.kw, cgv = q.cg / q.vec, nct = q.cg / q.cc, lpp = nct * q.ns, hv = q.vec >> 1;
    const int nft = (q.kw > 1) + (q.kh > 1), not_ = K - 1 - nft;
    const int NF = nft * cgv, NO = (not_ + 1) * cgv;
    const int nvf = (NF + q.ns - 1) / q.ns, nvo = (NO + q.ns - 1) / q.ns;
    const int pf = nvf * hv, po = nvo * hv;
    const int npairs = q.cc * (pf + po), nw4 = (npairs + 1) / 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");           // the canonical rows come from the launch(es) before
    if (e >= nw4 * lpp) return;
    const int f2 = e / lpp, l = e - f2 * lpp;                    // float4 number, lane
    const int G = blockIdx.y % q.groups, r = blockIdx.y / q.groups;
    const int dir = r & 1, layer = r >> 1;
    const int ks = l % q.ns, ct = l / q.ns;
    float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int f = 2 * f2 + h;                                // flat pair index (the pad pair of an odd count too)
        if (f >= npairs) continue;
        const int i = f / q.cc, cc = f - i * q.cc;               // i-th packed pair of the slice
        const bool fresh = i < pf;
        const int ip = fresh ? i : i - pf;
        const int j = hv == 1 ? ip : ip >> 1, e2 = hv == 1 ? 0 : ip & 1;     // vector entry slot, pair inside it
        const int ent = j * q.ns + ks;
        const bool valid = ent < (fresh ? NF : NO);
        int code = 0;
        if (valid) {
            const int ti = ent / cgv, qv = ent - ti * cgv;
            int t = 0;
            bool is_x = false;
            if (fresh) t = (q.kw > 1 && ti == 0) ? 1 : q.kw;
            else if (ti == not_) is_x = true;
            else t = q.old_tap[ti];
            const float *src = q.prepared + (size_t)layer * q.prepared_stride +
                               ((size_t)dir * q.C + (size_t)G * q.cg + ct * q.cc + cc) * q.KDP + t * q.cg +
                               qv * q.vec + 2 * e2;
            w[2 * h] = src[0];
            w[2 * h + 1] = src[1];
            const int th = t / q.kw;
            code = wave_code(th, t - th * q.kw, qv * q.vec, is_x);
        }
        // (padding entries: offset 0, zero weights)
        if (cc == 0 && e2 == 0 && dir == 0 && G == 0)
            reinterpret_cast<int *>(q.pack + (size_t)layer * q.pack_stride + q.pack_floats)[(size_t)((fresh ? 0 : nvf) + j) * lpp + l] = code;
    }
    float *dstp = q.pack + (size_t)layer * q.pack_stride + (((size_t)dir * q.groups + G) * nw4 + f2) * lpp * 4 + (size_t)l * 4;
    if (q.aligned) *reinterpret_cast<float4 *>(dstp) = make_float4(w[0], w[1], w[2], w[3]);
    else { dstp[0] = w[0]; dstp[1] = w[1]; dstp[2] = w[2]; dstp[3] = w[3]; }     // (such a buffer is refused by the solve)
}

// ---- host side -----------------------------------------------------------------------------
// X(CG, KH, KW, CC, NS, VEC, ITERS, NTHR)
#define IFK_WAVE_VARIANTS                                                                           \
    X(12, 3, 3, 6, 4, 2, 1, 128) X(12, 3, 3, 6, 4, 2, 2, 128) X(12, 3, 3, 6, 4, 2, 4, 128)           \
    X(12, 3, 3, 6, 8, 2, 1, 256) X(12, 3, 3, 6, 8, 2, 2, 256)                                        \
    X(24, 3, 3, 6, 8, 2, 1, 256) X(24, 3, 3, 6, 8, 2, 2, 256)                                        \
    X(24, 3, 3, 6, 16, 2, 2, 256) X(24, 3, 3, 6, 16, 2, 4, 256)                                      \
    X(48, 3, 3, 6, 16, 2, 2, 256)                                                                    \
    X(48, 3, 3, 6, 32, 2, 4, 256)                                                                    \
    X(6, 3, 3, 6, 4, 2, 1, 256) X(6, 3, 3, 6, 4, 2, 2, 256)                                          \
    X(12, 3, 3, 3, 2, 2, 1, 128) X(12, 3, 3, 3, 2, 2, 2, 128)                                        \
    X(24, 3, 3, 3, 4, 2, 1, 256) X(24, 3, 3, 3, 4, 2, 2, 256)                                        \
    X(48, 3, 3, 3, 8, 2, 2, 256)

struct WaveVariant {
    int cg, kh, kw, cc, ns, vec, iters, nthr;
};
static const WaveVariant kWaveVariants[] = {
#define X(CG, KHc, KWc, CC, NS, VEC, IT, NT) {CG, KHc, KWc, CC, NS, VEC, IT, NT},
    IFK_WAVE_VARIANTS
#undef X
};

struct WaveConfig {
    bool ok;
    WaveVariant v;
    int nslots, threads, PS, RSP, YN, XN, grid_x;
    size_t smem_bytes;
    double conflict;      // modelled shared-memory wavefronts per ideal wavefront of the gather (1.0 = none)
};

// Shared-memory wavefronts of one wavefront step's gathers for a (PS, RSP) layout: every load
// instruction is replayed per phase (8 lanes for 128-bit, 16 for 64-bit accesses) as many times as
// the most loaded bank has distinct words.  All rows live, mid-image diagonal.
static double wave_gather_cost(const WaveVariant &v, int H, int W, int nslots, int threads, int PS, int RSP)
{
    const int nct = v.cg / v.cc, lpp = nct * v.ns, cgv = v.cg / v.vec, K = v.kh * v.kw;
    const int nft = (v.kw > 1) + (v.kh > 1), not_ = K - 1 - nft;
    const int NF = nft * cgv, NO = (not_ + 1) * cgv;
    const int nvf = (NF + v.ns - 1) / v.ns, nvo = (NO + v.ns - 1) / v.ns;
    const int phase = v.vec == 4 ? 8 : 16;
    const int YN = (H + v.kh - 1) * RSP;
    long total = 0, ideal = 0;
    const int d = (H + W) / 2;
    for (int w0 = 0; w0 < threads; w0 += 32) {
        for (int j = 0; j < nvf + nvo; j++) {
            const bool fresh = j < nvf;
            for (int ph = 0; ph < 32; ph += phase) {
                int words[32][4];
                int nw = 0, bankcnt[32] = {0};
                for (int ln = ph; ln < ph + phase; ln++) {
                    const int tid = w0 + ln;
                    const int l = tid % lpp, slot = tid / lpp, ks = l % v.ns;
                    int row = slot, col = d - slot;
                    if (slot >= nslots || row >= H || col < 0 || col >= W) { row = 0; col = 0; }
                    if (!fresh) col += 1;
                    const int e = (fresh ? j : j - nvf) * v.ns + ks;
                    int off = 0;
                    if (fresh) {
                        if (e < NF) {
                            const int ft = e / cgv, q = e % cgv;
                            const int t = (v.kw > 1 && ft == 0) ? 1 : v.kw;
                            off = -(t / v.kw) * RSP - (t % v.kw) * PS + q * v.vec;
                        }
                    } else if (e < NO) {
                        const int oi = e / cgv, q = e % cgv;
                        int t = 0;
                        if (oi < not_) {
                            int n = 0;
                            for (int tt = 1; tt < K; tt++) {
                                if (tt / v.kw + tt % v.kw < 2) continue;
                                if (n == oi) { t = tt; break; }
                                n++;
                            }
                        }
                        off = -(t / v.kw) * RSP - (t % v.kw) * PS + q * v.vec + (oi == not_ ? YN : 0);
                    }
                    const int base = (row + v.kh - 1) * RSP + (col + v.kw - 1) * PS + off;
                    for (int e2 = 0; e2 < v.vec; e2++) {
                        const int word = base + e2;
                        bool seen = false;
                        for (int k = 0; k < nw; k++) seen |= words[k / 4][k % 4] == word;
                        if (!seen && nw < 128) {
                            words[nw / 4][nw % 4] = word;
                            nw++;
                            bankcnt[((word % 32) + 32) % 32]++;
                        }
                    }
                }
                int worst = 1;
                for (int bk = 0; bk < 32; bk++) worst = bankcnt[bk] > worst ? bankcnt[bk] : worst;
                total += worst;
                ideal += 1;
            }
        }
    }
    return ideal ? (double)total / ideal : 1.0;
}

// (cc, ns, vec) family serving a (Cg, KH, KW): fixes the packed-weight layout, so it must not depend on
// the image size or the batch -- every variant of one (Cg, KH, KW) in the table shares it
static const WaveVariant *wave_family(const Geometry &g)
{
    const EnvKnobs &k = env();
    for (const WaveVariant &v : kWaveVariants) {
        if (v.cg != g.Cg || v.kh != g.KH || v.kw != g.KW) continue;
        if (k.wave_cfg[0] && (k.wave_cfg[0] != v.cc || k.wave_cfg[1] != v.ns || k.wave_cfg[2] != v.vec)) continue;
        return &v;
    }
    return nullptr;
}

static WaveConfig choose_wave_uncached(const Geometry &g)
{
    WaveConfig c{};
    c.ok = false;
    const EnvKnobs &k = env();
    if (k.wave_off || k.pins_other_solver()) return c;
    const int max_smem = device_max_smem_optin();
    const WaveVariant *fam = wave_family(g);
    if (!fam) return c;
    for (const WaveVariant &v : kWaveVariants) {
        if (v.cg != g.Cg || v.kh != g.KH || v.kw != g.KW) continue;
        if (g.W < 2 || g.H * g.W < 2) continue;               // the index divisions use 32-bit magic multipliers
        if (v.cc != fam->cc || v.ns != fam->ns || v.vec != fam->vec) continue;
        const int lpp = (v.cg / v.cc) * v.ns;
        int nslots = v.nthr / lpp;
        if (nslots < 1) continue;
        if (nslots > g.H) nslots = g.H;
        if (nslots * v.iters < g.H) continue;                 // rows per thread of this variant do not cover H
        const int threads = round_up(nslots * lpp, 32);
        // layout: odd pixel stride in vector units; row pad by enumeration of the gather's bank conflicts
        int PS = round_up(g.Cg, v.vec);
        if (((PS / v.vec) & 1) == 0) PS += v.vec;
        const int row_px = g.W + g.KW;                        // halo left + one look-ahead column right
        int best_pad = 0;
        double best_cost = 1e30;
        for (int pad = 0; pad < 32; pad += v.vec) {
            const double cost = wave_gather_cost(v, g.H, g.W, nslots, threads, PS, row_px * PS + pad);
            if (cost < best_cost - 1e-9) { best_cost = cost; best_pad = pad; }
        }
        const int RSP = row_px * PS + best_pad;
        const int YN = round_up((g.H + g.KH - 1) * RSP, 4);
        const int XN = round_up(g.Cg * g.H * g.W, 4);
        const size_t smem = 16 + ((size_t)XN + 2 * (size_t)YN) * sizeof(float);
        if (smem > (size_t)max_smem) continue;
        c.ok = true;
        c.v = v; c.nslots = nslots; c.threads = threads; c.PS = PS; c.RSP = RSP; c.YN = YN; c.XN = XN;
        c.smem_bytes = smem; c.conflict = best_cost;
        break;                                                // variants are listed by rising rows per thread
    }
    if (!c.ok) return c;
    // one CTA per SM is all the register file holds; a stripe of images per CTA beyond that
    int grid_x = (device_sm_count() + g.groups - 1) / g.groups;
    if (grid_x > g.B) grid_x = g.B;
    if (grid_x < 1) grid_x = 1;
    c.grid_x = grid_x;
    return c;
}

static WaveConfig choose_wave(const Geometry &g)
{
    // memoised per geometry (the enumeration above costs tens of microseconds; a launch must not)
    typedef std::tuple<int, int, int, int, int, int, int> Key;
    static std::map<Key, WaveConfig> cache;
    static std::mutex mu;
    static unsigned seen_generation = 0;
    const unsigned gen = env_generation();
    const Key key(g.C, g.H, g.W, g.KH, g.KW, g.groups, g.B < device_sm_count() ? g.B : device_sm_count());
    std::lock_guard<std::mutex> lock(mu);
    if (seen_generation != gen) {                                     // knobs reloaded (tests): start over
        cache.clear();
        seen_generation = gen;
    }
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const WaveConfig c = choose_wave_uncached(g);
    cache[key] = c;
    return c;
}

bool wave_solve_available(const Geometry &g) { return choose_wave(g).ok; }

int describe_wave_solve(const Geometry &g, char *buf, size_t buflen)
{
    const WaveConfig c = choose_wave(g);
    snprintf(buf, buflen, "wave<cg=%d,k=%dx%d,cc=%d,ns=%d,vec=%d,iters=%d> slots=%d threads=%d ps=%d rsp=%d conflict=%.2f "
             "smem=%zuB grid=%dx%d", c.v.cg, c.v.kh, c.v.kw, c.v.cc, c.v.ns, c.v.vec, c.v.iters, c.nslots, c.threads,
             c.PS, c.RSP, c.conflict, c.smem_bytes, c.grid_x, g.groups);
    return 0;
}

struct WavePackDims {
    int lpp, nvf, nvo, nw4;
    size_t pack_floats, code_ints;      // per layer: both directions / all groups; codes once
};
static WavePackDims wave_pack_dims(const WaveVariant &v, int groups)
{
    WavePackDims d{};
    const int K = v.kh * v.kw, cgv = v.cg / v.vec, nct = v.cg / v.cc;
    const int nft = (v.kw > 1) + (v.kh > 1), not_ = K - 1 - nft;
    d.lpp = nct * v.ns;
    d.nvf = (nft * cgv + v.ns - 1) / v.ns;
    d.nvo = ((not_ + 1) * cgv + v.ns - 1) / v.ns;
    d.nw4 = (v.cc * (d.nvf + d.nvo) * v.vec / 2 + 1) / 2;
    d.pack_floats = (size_t)2 * groups * d.nw4 * d.lpp * 4;
    d.code_ints = (size_t)round_up((d.nvf + d.nvo) * d.lpp, 4);
    return d;
}

size_t prepared_floats(const Geometry &g)
{
    size_t n = (size_t)2 * g.C * g.KDP;                       // canonical rows, both directions
    if (const WaveVariant *v = wave_family(g)) {
        const WavePackDims d = wave_pack_dims(*v, g.groups);
        n += d.pack_floats + d.code_ints;
    }
    return n + split_pack_floats(g);                          // the split kernel's packed copy comes last
}

size_t split_pack_offset(const Geometry &g) { return prepared_floats(g) - split_pack_floats(g); }

int launch_wave_pack(const Geometry &g, float *prepared, int count, size_t prepared_stride, cudaStream_t s)
{
    if (count <= 0) return 0;
    if (int st = launch_split_pack(g, prepared, prepared + split_pack_offset(g), count, prepared_stride, s)) return st;
    const WaveVariant *v = wave_family(g);
    if (!v) return 0;
    const WavePackDims d = wave_pack_dims(*v, g.groups);
    WavePackParams q{};
    q.prepared = prepared;
    q.pack = prepared + (size_t)2 * g.C * g.KDP;
    q.pack_floats = d.pack_floats;
    q.prepared_stride = prepared_stride;
    q.pack_stride = prepared_stride;
    q.C = g.C; q.cg = v->cg; q.kh = v->kh; q.kw = v->kw; q.cc = v->cc; q.ns = v->ns; q.vec = v->vec;
    q.KDP = g.KDP; q.groups = g.groups; q.count = count;
    {   // taps two or more diagonals back, in array order (wave_gather_cost enumerates them the same way)
        int n = 0;
        for (int tt = 1; tt < g.K && n < 64; tt++)
            if (tt / v->kw + tt % v->kw >= 2) q.old_tap[n++] = (unsigned char)tt;
    }
    if (g.K > 64 || (long)count * 2 * g.groups > 65535) return IFK_ERR_UNSUPPORTED;
    q.aligned = ((uintptr_t)q.pack % 16 == 0 && (count == 1 || prepared_stride % 4 == 0)) ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((d.nw4 * d.lpp + 255) / 256), (unsigned)(count * 2 * g.groups));
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // scheduled while the prepare kernels still run
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env().pdl ? 1 : 0;
    return cuda_status(cudaLaunchKernelEx(&cfg, wave_pack_kernel, q));
}

// one launch over `n` consecutive layers (n == 1: a plain solve).  prepared[i]: layer i's whole prepared buffer.
static int launch_wave_layers(const Geometry &g, int n, const int *orients, const float *const *prepared, const float *in,
                              float *const *outs, bool reverse, int flags, long long *probe, cudaStream_t s,
                              const ifk_fused *fused = nullptr, float *out2 = nullptr)
{
    const WaveConfig c = choose_wave(g);
    if (!c.ok || n < 1 || n > kWaveChainMax) return IFK_ERR_UNSUPPORTED;
    if (fused && (n != 1 || c.v.vec != 2)) return IFK_ERR_UNSUPPORTED;
    const WavePackDims d = wave_pack_dims(c.v, g.groups);
    WaveParams p{};
    p.in = in;
    p.nlayers = n;
    bool aligned = true;
    for (int i = 0; i < n; i++) {
        const float *pack = prepared[i] + (size_t)2 * g.C * g.KDP;
        p.pack[i] = reinterpret_cast<const float4 *>(pack + (reverse ? d.pack_floats / 2 : 0));
        p.out[i] = outs[i];
        const int orient = orients ? orients[i] : g.orient;
        p.flips[i] = reverse ? (orient ^ 3) : orient;      // the adjoint walks the fully reflected frame
        aligned = aligned && ((uintptr_t)prepared[i] % 16 == 0);
    }
    if (!aligned) return IFK_ERR_UNSUPPORTED;               // the packed weights are read as 16-byte words
    p.codes = reinterpret_cast<const int *>(prepared[0] + (size_t)2 * g.C * g.KDP + d.pack_floats);
    p.B = g.B; p.C = g.C; p.H = g.H; p.W = g.W;
    p.nslots = c.nslots;
    const size_t img_bytes = (size_t)g.Cg * g.H * g.W * sizeof(float);
    p.bulk = (img_bytes % 16 == 0) && ((uintptr_t)in % 16 == 0) && !env().nobulk ? 1 : 0;
    p.early = (flags & IFK_FLAG_STABLE_PREPARED) ? 1 : 0;
    p.PS = c.PS; p.RSP = c.RSP; p.YN = c.YN; p.XN = c.XN;
    p.mW = (unsigned)((0x100000000ULL + (unsigned)g.W - 1) / (unsigned)g.W);
    {   // pixel lanes: the largest power of two <= H*W that divides the thread count
        int sh = 0;
        while ((2 << sh) <= g.H * g.W && c.threads % (2 << sh) == 0) sh++;
        p.tm_shift = sh;
    }
    p.probe = probe;
    if (fused) {
        p.in_scale = reverse ? nullptr : fused->in_scale;
        p.in_bias = reverse ? nullptr : fused->in_bias;
        p.out_scale = reverse ? fused->out_scale : nullptr;
        p.out2 = reverse ? out2 : nullptr;
        p.squeeze_in = reverse ? 0 : fused->squeeze;
        p.squeeze_out = reverse ? fused->squeeze : 0;
        p.fused = 1;
        if (fused->squeeze && (((uintptr_t)in | (uintptr_t)out2) % 8 != 0)) return IFK_ERR_UNSUPPORTED;   // channel pairs move as float2
    }
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(c.threads);
    cfg.dynamicSmemBytes = c.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL, see the kernel prologue
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env().pdl ? 1 : 0;
#define X(CG, KHc, KWc, CC, NS, VEC, IT, NT)                                                          \
    if (c.v.cg == CG && c.v.kh == KHc && c.v.kw == KWc && c.v.cc == CC && c.v.ns == NS && c.v.vec == VEC && \
        c.v.iters == IT && c.v.nthr == NT) {                                                          \
        auto kern = solve_wave_kernel<CG, KHc, KWc, CC, NS, VEC, IT, NT>;                             \
        if (c.smem_bytes > 48 * 1024) {                                                               \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 (int)c.smem_bytes);                                  \
            if (e != cudaSuccess) return (int)e;                                                      \
        }                                                                                             \
        /* CTAs that fit one SM at once (registers, shared memory): a batch beyond the SM count runs that */ \
        /* many images per SM concurrently, whose latency-bound wavefronts interleave */               \
        int occ = 1;                                                                                  \
        if (g.B * g.groups > device_sm_count() &&                                                     \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, c.threads, c.smem_bytes) != cudaSuccess) \
            occ = 1;                                                                                  \
        if (occ < 1) occ = 1;                                                                         \
        int gx = (device_sm_count() * occ + g.groups - 1) / g.groups;                                 \
        if (gx > g.B) gx = g.B;                                                                       \
        cfg.gridDim = dim3(gx < 1 ? 1 : gx, g.groups);                                                \
        return cuda_status(cudaLaunchKernelEx(&cfg, kern, p));                                        \
    }
    IFK_WAVE_VARIANTS
#undef X
    return IFK_ERR_UNSUPPORTED;
}

int launch_solve_wave(const Geometry &g, const float *in, const float *prepared, float *out, bool reverse,
                      int flags, long long *probe, cudaStream_t s)
{
    return launch_wave_layers(g, 1, nullptr, &prepared, in, &out, reverse, flags, probe, s);
}

int launch_solve_wave_fused(const Geometry &g, const ifk_fused &f, const float *in, const float *prepared, float *out,
                            float *out2, bool reverse, cudaStream_t s)
{
    return launch_wave_layers(g, 1, nullptr, &prepared, in, &out, reverse, g.flags, nullptr, s, &f, out2);
}

// consecutive layers feeding each other (ifk_inverse_chain_f32): groups of up to kWaveChainMax layers per launch
int launch_solve_chain(const Geometry &g, int n, const int *orients, const float *const *prepared, const float *x,
                       float *const *ys, cudaStream_t s)
{
    if (g.B == 0) return 0;
    if (split_solve_available(g)) {
        const size_t off = split_pack_offset(g);
        const float *in = x;
        for (int i = 0; i < n; i += kWaveChainMax) {
            const int m = n - i < kWaveChainMax ? n - i : kWaveChainMax;
            const float *packs[kWaveChainMax];
            for (int k = 0; k < m; k++) packs[k] = prepared[i + k] + off;
            const int st = launch_split_layers(g, m, orients + i, packs, in, ys + i, false, 0, nullptr, s);
            if (st != 0) return st;
            in = ys[i + m - 1];
        }
        return 0;
    }
    if (!choose_wave(g).ok) return IFK_ERR_UNSUPPORTED;
    const float *in = x;
    for (int i = 0; i < n; i += kWaveChainMax) {
        const int m = n - i < kWaveChainMax ? n - i : kWaveChainMax;
        // (no IFK_FLAG_STABLE_PREPARED: the caller may have prepared the weights right before)
        const int st = launch_wave_layers(g, m, orients + i, prepared + i, in, ys + i, false, 0, nullptr, s);
        if (st != 0) return st;
        in = ys[i + m - 1];
    }
    return 0;
}

}  // namespace ifk
