// Shared-memory resident wavefront solve kernel (template) and its parameter block.
// Included by the per-vector-width translation units ifk_solve_v{1,2,4}.cu.
//
// Replaces the reference's per-diagonal launch loop
// (inf/utils/inv_conv_cuda/inv_conv_with_bp_kernel_general.cu:72-129: (H+W-1)*C/4 launches,
// each followed by cudaDeviceSynchronize, one thread per (batch, group, pixel), dependent
// global read-modify-writes) by ONE launch.  Per image of the CTA's batch stripe:
//   1. the group's image (contiguous in NCHW) lands in `xbuf` by one TMA bulk copy
//      (cp.async.bulk + mbarrier); the next image of the stripe is prefetched as soon as
//      xbuf has been consumed;
//   2. pre-pass, no dependencies: z = T x for every pixel into `zbuf` (T = (I + A0)^-1 is
//      tap 0 of the prepared kernel; skipped when Cg == 1);
//   3. wavefront: thread -> (row slot, ct, ks) walks its image row, one pixel per
//      anti-diagonal.  `ct` = tile of CC output channels, `ks` = slice of the (K-1)*Cg
//      neighbour reduction.  y lives in `ybuf` as [row][col][channel] with a zero halo, so a
//      neighbour pixel's channels are VEC-wide contiguous vectors: NV vector loads per thread
//      per pixel, weights of the slice in registers for the whole stripe, partial sums
//      combined by full-warp shuffles, the ks == 0 lane adds z and writes y into `ybuf`
//      (for later diagonals) and in place into `zbuf`; one block barrier per diagonal;
//   4. `zbuf` returns to global memory by one TMA bulk store.
// The adjoint solve is the same walk in reflected coordinates, and so are the TR/BL/BR
// orientations (ifk.h): only the index into the contiguous buffers is mirrored (`flip`).
#pragma once
#include "ifk_internal.cuh"

namespace ifk {

struct SolveParams {
    const float *in;
    float *out;
    const float *prep;  // prepared weights of this direction: [group][co][KDP]
    int B, C, H, W, KH, KW, Cg, KD, KDP;
    int WP;             // halo-padded width  (W + KW - 1)
    int PS;             // pixel stride of ybuf in floats (channels rounded up, conflict padded)
    int YN;             // floats in ybuf = (H + KH - 1) * WP * PS
    int XN;             // floats per contiguous image buffer (Cg*H*W rounded up to 4)
    int CgV;            // channel vectors per pixel = ceil(Cg / VEC)
    int NVT;            // (K - 1) * CgV : vector entries of one output row's reduction
    int CgP4;           // Cg rounded up to 4 (row stride of the transposed T in smem)
    int kw_magic;       // ceil(65536 / KW): t / KW == (t * kw_magic) >> 16 for every tap index
    int v_dt, v_dq;     // NS / CgV and NS % CgV: how a thread's vector index advances per j
    int NS, NCT, nslots, iters;
    int nwork;          // threads that walk the wavefront (multiple of 32); the rest only help staging
    int flip;           // memory index of solver pixel (h, w): bit 0 -> column W-1-w, bit 1 -> row H-1-h
                        // (adjoint solve = both reflected; IFK_ORIENT_* reflections are XORed in)
    int bulk;           // image size / pointers allow TMA bulk copies (16-byte granularity)
    int walign;         // prepared weights are 16-byte aligned (128-bit weight loads allowed)
    int early;          // IFK_FLAG_STABLE_PREPARED: weights may be fetched ahead of griddepcontrol.wait
    long long *probe;   // tuning aid: clock64() stamps of CTA (0,0) thread 0, or nullptr
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

// Pin loop-invariant values in registers.  Left alone, ptxas re-reads kernel parameters from the
// constant bank inside the diagonal loop (rematerialisation), and those loads -- tens of cycles
// each -- sit on the loop-carried critical path of a latency-bound kernel.  Adding a zero that
// was read back from shared memory through a volatile load makes the value opaque.
struct Hold {
    int zero;
    __device__ __forceinline__ int operator()(int v) const { return v + zero; }
    __device__ __forceinline__ uint32_t operator()(uint32_t v) const { return v + (uint32_t)zero; }
};

// ---- vector loads from shared memory by 32-bit shared-space address -------------------------
template <int VEC>
__device__ __forceinline__ void lds_vec(float *v, uint32_t addr);
template <>
__device__ __forceinline__ void lds_vec<1>(float *v, uint32_t addr)
{
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr) : "memory");
}
template <>
__device__ __forceinline__ void lds_vec<2>(float *v, uint32_t addr)
{
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr) : "memory");
}
template <>
__device__ __forceinline__ void lds_vec<4>(float *v, uint32_t addr)
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                 : "r"(addr)
                 : "memory");
}

// ---- packed FP32: two FMAs per instruction (SASS FFMA2) ---------------------------------------
// A three-register FFMA issues every other cycle per scheduler on this architecture; the packed form
// does two per instruction, so FMA-bound inner loops need half the issue slots.  The pairs run along
// the REDUCTION axis -- (w[i], w[i+1]) * (v[i], v[i+1]) accumulate into (lo, hi) halves that are added
// at the end -- so loaded values pair up as they arrive (ld.shared.v2.b64) and nothing is duplicated.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack_f32x2(float lo, float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float sum_f32x2(f32x2_t v)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ f32x2_t fma_f32x2(f32x2_t a, f32x2_t b, f32x2_t c)
{
    f32x2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// VEC floats from shared memory as VEC/2 packed pairs
template <int VEC>
__device__ __forceinline__ void lds_pairs(f32x2_t *v, uint32_t addr);
template <>
__device__ __forceinline__ void lds_pairs<2>(f32x2_t *v, uint32_t addr)
{
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v[0]) : "r"(addr) : "memory");
}
template <>
__device__ __forceinline__ void lds_pairs<4>(f32x2_t *v, uint32_t addr)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "r"(addr) : "memory");
}

// ---- reduce-scatter of N partial sums over the NS adjacent lanes of a pixel ------------------
// Recursive halving: at the level with lane mask m every lane keeps one half of its current
// values and receives the partner's partial sums of that half (N/2 + N/4 + ... shuffles in
// total, against N*log2(NS) for a butterfly) and the finished channels end up spread over the
// lanes, which then write them.  The value counts per level are compile-time; NS is a runtime
// (warp-uniform) parameter that only decides how many levels run.  After the last level lane
// `ks` holds the complete sums of channels [off, off + size) of its tile in acc[0 .. size),
// (off, size) = rs_owner(CC, NS, ks).
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <int N, int LEVELS>
struct Rs {
    // zv: z of the channels this lane finishes; ya/za: shared addresses of their y / in-place slots
    __device__ __forceinline__ static void run(float *acc, const float *zv, int ks, int m, int own_size,
                                               bool active, uint32_t ya, uint32_t za, uint32_t zstride)
    {
        if (m == 0 || LEVELS == 0) {
#pragma unroll
            for (int i = 0; i < N; i++)
                if (active && i < own_size) {
                    const float yv = acc[i] + zv[i];
                    sts_f32(ya + 4u * i, yv);
                    sts_f32(za + zstride * i, yv);
                }
            return;
        }
        constexpr int HALF = (N + 1) / 2;
        const bool hi = (ks & m) != 0;
#pragma unroll
        for (int i = 0; i < HALF; i++) {
            const float lo_v = acc[i];
            const float hi_v = (i + HALF < N) ? acc[i + HALF] : 0.f;
            const float send = hi ? lo_v : hi_v;
            const float keep = hi ? hi_v : lo_v;
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
        Rs<HALF, (LEVELS > 0 ? LEVELS - 1 : 0)>::run(acc, zv, ks, m >> 1, own_size, active, ya, za, zstride);
    }
};

__device__ __forceinline__ void rs_owner(int n, int m, int ks, int *off, int *size)
{
    int o = 0, sz = n;
    for (; m > 1; m >>= 1) {
        const int half = (n + 1) / 2;
        if (ks & (m / 2)) { o += half; sz = sz - half > 0 ? sz - half : 0; }
        else              { sz = sz < half ? sz : half; }
        n = half;
    }
    *off = o;
    *size = sz;
}

template <int CC, int NV, int VEC>
constexpr int solve_max_threads()
{
    // registers: CC*NV*VEC weights + NV*VEC loaded values + NV offsets + CC sums + ~48 of bookkeeping
    int regs = CC * NV * VEC + NV * VEC + NV + 2 * CC + 48;
    if (regs > 255) regs = 255;
    int t = (65536 / regs) / 32 * 32;
    return t > 1024 ? 1024 : t;
}

#define IFK_PROBE(i) do { if (p.probe && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.probe[i] = clock64(); } while (0)

template <int CC, int NV, int VEC>
__global__ void __launch_bounds__(solve_max_threads<CC, NV, VEC>())
solve_smem_kernel(const SolveParams p)
{
    IFK_PROBE(0);
    extern __shared__ __align__(128) float smem[];
    const int Cg = p.Cg, WP = p.WP, PS = p.PS, H = p.H, W = p.W, HW = p.H * p.W;
    const int tid = threadIdx.x, nthr = blockDim.x;

    // shared memory carve-up (every region a multiple of 16 bytes)
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);         // 16 bytes reserved
    float *tT = smem + 4;                                       // [Cg][CgP4] transposed T
    float *xbuf = tT + (Cg > 1 ? Cg * p.CgP4 : 0);              // [Cg][HW] raw input
    float *zbuf = Cg > 1 ? xbuf + p.XN : xbuf;                  // [Cg][HW] T x, then y in place
    float *ybuf = zbuf + p.XN;                                  // [HP][WP][PS] y, zero halo top/left

    const int G = blockIdx.y;
    const float *wg = p.prep + (size_t)G * Cg * p.KDP;
    const uint32_t img_bytes = (uint32_t)(Cg * HW) * 4u;
    const size_t img_stride = (size_t)p.C * HW;
    const float *in0 = p.in + (size_t)G * Cg * HW;
    float *out0 = p.out + (size_t)G * Cg * HW;

    // Programmatic dependent launch: everything that does not depend on the previous kernel's output
    // happens before the dependency wait: the shared-memory zero fill and -- only when the caller
    // vouches that the previous operation of the stream did not write the prepared weights
    // (IFK_FLAG_STABLE_PREPARED, ifk.h) -- fetching this thread's weight slice and T.  Without the flag
    // the fetch follows the wait: CUDA guarantees a predecessor's writes only after it.  The dependents
    // are released after the wait, so the guarantee is transitive along a chain of solves.
    if (!p.early) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (tid == 0) smem[2] = 0.f;          // source of the opaque zero used by Hold (see above)

    // ybuf starts from zero for every image: the halo, and the not-yet-written interior that
    // zero-weight padding entries may touch
    for (int i = tid * 4; i < p.YN; i += nthr * 4)
        *reinterpret_cast<float4 *>(ybuf + i) = make_float4(0.f, 0.f, 0.f, 0.f);

    int b = blockIdx.x;
    const int NS = p.NS, NCT = p.NCT;
    const int ks = tid % NS;
    const int ct = (tid / NS) % NCT;
    const int slot = tid / (NS * NCT);
    const bool worker = slot < p.nslots;

    // this thread's slice of the prepared kernel -> registers, for the whole batch stripe
    float wreg[CC][NV * VEC];
    int offs[NV];
    // vector entry v = j*NS + ks -> (tap t1+1, channel vector q); advanced without divisions
    int t1 = ks / p.CgV, q = ks - t1 * p.CgV;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const bool valid = worker && j * NS + ks < p.NVT;
        const int t = t1 + 1;
        const int qh = (t * p.kw_magic) >> 16, qw = t - qh * p.KW;
        offs[j] = valid ? ((-qh * WP - qw) * PS + q * VEC) * 4 : 0;   // padding entries: weight 0, finite data
        const int wcol = valid ? t1 * Cg + q * VEC : 0;               // weight column of its 1st channel
        const int ci0 = valid ? q * VEC : 0;
        q += p.v_dq;
        t1 += p.v_dt;
        if (q >= p.CgV) { q -= p.CgV; t1++; }
        if (VEC == 4 && (Cg & 3) == 0 && p.walign) {
            // 4 consecutive input channels of one tap: one 128-bit load per output channel
#pragma unroll
            for (int cc = 0; cc < CC; cc++) {
                const int co = ct * CC + cc;
                float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && co < Cg) w4 = __ldg(reinterpret_cast<const float4 *>(wg + (size_t)co * p.KDP + Cg + wcol));
                wreg[cc][j * VEC + 0] = w4.x;
                wreg[cc][j * VEC + 1 % VEC] = w4.y;
                wreg[cc][j * VEC + 2 % VEC] = w4.z;
                wreg[cc][j * VEC + 3 % VEC] = w4.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < VEC; e++)
#pragma unroll
                for (int cc = 0; cc < CC; cc++) {
                    const int co = ct * CC + cc;
                    wreg[cc][j * VEC + e] = (valid && co < Cg && ci0 + e < Cg)
                                                ? __ldg(wg + (size_t)co * p.KDP + Cg + wcol + e) : 0.f;
                }
        }
    }

    // the same weights as packed pairs along the reduction axis (the scalar array is dead afterwards)
    f32x2_t w2[CC][(NV * VEC + 1) / 2];
    if constexpr (VEC >= 2) {
#pragma unroll
        for (int cc = 0; cc < CC; cc++)
#pragma unroll
            for (int i = 0; i < NV * VEC / 2; i++) w2[cc][i] = pack_f32x2(wreg[cc][2 * i], wreg[cc][2 * i + 1]);
    }

    if (Cg > 1)
        for (int i = tid; i < Cg * p.CgP4; i += nthr) {
            const int ci = i / p.CgP4, co = i - ci * p.CgP4;
            tT[i] = co < Cg ? __ldg(wg + (size_t)co * p.KDP + ci) : 0.f;
        }
    if (p.early) {                                          // the input image may only be touched from here on
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (p.bulk && tid == 0) {
        mbar_init(bar, 1);
        if (b < p.B) {
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b * img_stride, img_bytes, bar);
        }
    }
    __syncthreads();        // T, mbarrier init, Hold's zero visible
    IFK_PROBE(1);

    // which of the tile's CC output channels this lane finishes after the reduce-scatter
    int own_off, own_size;
    rs_owner(CC, NS, ks, &own_off, &own_size);
    {
        const int tile_n = Cg - ct * CC < CC ? Cg - ct * CC : CC;     // channels in this (last) tile
        own_size = own_off + own_size > tile_n ? (tile_n - own_off > 0 ? tile_n - own_off : 0) : own_size;
    }
    const int own_c0 = ct * CC + own_off;

    IFK_PROBE(2);
    Hold hold;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hold.zero) : "r"(smem_u32(smem + 2)) : "memory");
    const int ndiag = hold(H + W - 1);
    const int iters = hold(p.iters), nslots = hold(p.nslots), nwork = hold(p.nwork), Wr = hold(W);
    const int NSr = hold(NS >> 1);
    own_size = hold(own_size);
    const uint32_t ybase = hold(smem_u32(ybuf) + (uint32_t)(((p.KH - 1) * WP + (p.KW - 1)) * PS) * 4u);
    const uint32_t zstride = hold((uint32_t)HW * 4u);
    // rows of this thread: slot, slot + nslots, ...; row_ok = how many of them exist
    const int row_ok = hold(worker && slot < H ? (H - 1 - slot) / nslots + 1 : 0);
    // pixel (h, w = d - h): ybuf address = ybase + ((h*WP + w)*PS)*4 ; contiguous index rr = h*W + w
    const uint32_t pix_step = hold((uint32_t)PS * 4u);                                   // per diagonal
    const uint32_t pix_row = hold((uint32_t)(nslots * (WP - 1) * PS) * 4u);              // per row iteration
    const uint32_t pix0 = ybase + (uint32_t)(slot * (WP - 1) * PS) * 4u;                 // d = 0, it = 0
    // contiguous index of solver pixel (h, w) = idx0 + sh*h*W + sw*w  (sh, sw = -1 on a reflected axis)
    const int sw = (p.flip & 1) ? -1 : 1, sh = (p.flip & 2) ? -1 : 1;
    const int idx0 = ((p.flip & 2) ? (H - 1) * W : 0) + ((p.flip & 1) ? W - 1 : 0);
    const uint32_t z_step = hold((uint32_t)(sw * 4));
    const uint32_t z_row = hold((uint32_t)(nslots * (sh * W - sw) * 4));
    const uint32_t z0 = smem_u32(zbuf) + (uint32_t)(own_c0 * HW) * 4u + (uint32_t)((idx0 + slot * (sh * W - sw)) * 4);
    const uint32_t own_c0_bytes = hold((uint32_t)own_c0 * 4u);
    const int slot_r = hold(slot);
    uint32_t parity = 0;
    for (; b < p.B; b += gridDim.x) {
        const int b_next = b + gridDim.x;
        IFK_PROBE(3);
        if (p.bulk) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            const float *src = in0 + (size_t)b * img_stride;
            for (int i = tid; i < Cg * HW; i += nthr) xbuf[i] = __ldg(src + i);
        }
        IFK_PROBE(4);
        if (Cg > 1) {
            if (tid == 0 && p.bulk) bulk_store_wait_read();     // previous image has left zbuf
            __syncthreads();
            // pre-pass z = T x : item (pixel r, 4 output channels), consecutive threads -> pixels
            const int n4 = p.CgP4 >> 2;
            for (int i = tid; i < HW * n4; i += nthr) {
                const int c4 = i / HW, r = i - c4 * HW;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                const float *tp = tT + c4 * 4;
#pragma unroll 4
                for (int ci = 0; ci < Cg; ci++) {
                    const float xv = xbuf[ci * HW + r];
                    const float4 t4 = *reinterpret_cast<const float4 *>(tp + ci * p.CgP4);
                    a0 = fmaf(t4.x, xv, a0);
                    a1 = fmaf(t4.y, xv, a1);
                    a2 = fmaf(t4.z, xv, a2);
                    a3 = fmaf(t4.w, xv, a3);
                }
                const int co = c4 * 4;
                zbuf[co * HW + r] = a0;
                if (co + 1 < Cg) zbuf[(co + 1) * HW + r] = a1;
                if (co + 2 < Cg) zbuf[(co + 2) * HW + r] = a2;
                if (co + 3 < Cg) zbuf[(co + 3) * HW + r] = a3;
            }
        }
        __syncthreads();
        if (p.bulk && tid == 0 && Cg > 1 && b_next < p.B) {        // prefetch the next image
            mbar_expect_tx(bar, img_bytes);
            bulk_load(xbuf, in0 + (size_t)b_next * img_stride, img_bytes, bar);
        }

        // wavefront.  Row `slot + it*nslots` of this thread meets diagonal d at column d - row; all
        // addresses advance by a constant per diagonal, so a step is loads, FMAs, shuffles and
        // stores.  Branches are what an otherwise empty step costs (~150 cycles with the generic
        // loop nest, measured), hence the three specialised loops below.
        IFK_PROBE(5);
        auto step = [&](uint32_t pix, uint32_t za, bool active) {
            const uint32_t pa = active ? pix : ybase;              // idle lanes: a legal pixel
            if constexpr (VEC >= 2) {
                // packed path: values arrive as pairs, two FMAs per instruction
                f32x2_t v2[NV * VEC / 2];
#pragma unroll
                for (int j = 0; j < NV; j++) lds_pairs<VEC>(v2 + j * (VEC / 2), pa + (uint32_t)offs[j]);
                float zv[CC];
#pragma unroll
                for (int i = 0; i < CC; i++) zv[i] = (active && i < own_size) ? lds_f32(za + zstride * i) : 0.f;
                f32x2_t acc2[CC];
#pragma unroll
                for (int cc = 0; cc < CC; cc++) acc2[cc] = 0ull;
#pragma unroll
                for (int i = 0; i < NV * VEC / 2; i++)
#pragma unroll
                    for (int cc = 0; cc < CC; cc++) acc2[cc] = fma_f32x2(w2[cc][i], v2[i], acc2[cc]);
                float acc[CC];
#pragma unroll
                for (int cc = 0; cc < CC; cc++) acc[cc] = sum_f32x2(acc2[cc]);
                Rs<CC, 5>::run(acc, zv, ks, NSr, own_size, active, pa + own_c0_bytes, za, zstride);
                return;
            }
            float v[NV * VEC];
#pragma unroll
            for (int j = 0; j < NV; j++) lds_vec<VEC>(v + j * VEC, pa + (uint32_t)offs[j]);
            float zv[CC];
#pragma unroll
            for (int i = 0; i < CC; i++) zv[i] = (active && i < own_size) ? lds_f32(za + zstride * i) : 0.f;

            constexpr int NACC = CC >= 4 ? 1 : (CC >= 2 ? 2 : 4);  // independent FMA chains
            float part[NACC][CC];
#pragma unroll
            for (int a = 0; a < NACC; a++)
#pragma unroll
                for (int cc = 0; cc < CC; cc++) part[a][cc] = 0.f;
#pragma unroll
            for (int i = 0; i < NV * VEC; i++)
#pragma unroll
                for (int cc = 0; cc < CC; cc++) part[i % NACC][cc] = fmaf(wreg[cc][i], v[i], part[i % NACC][cc]);
            float acc[CC];
#pragma unroll
            for (int cc = 0; cc < CC; cc++) {
                acc[cc] = part[0][cc];
#pragma unroll
                for (int a = 1; a < NACC; a++) acc[cc] += part[a][cc];
            }
            Rs<CC, 5>::run(acc, zv, ks, NSr, own_size, active, pa + own_c0_bytes, za, zstride);
        };

        if (tid < nwork) {                       // helper warps skip the wavefront altogether
            uint32_t pix_d = pix0, z_d = z0;
            if (iters == 1 && nwork <= 32) {
                // one warp, one row per thread: every diagonal has a live row, no block barrier
                const unsigned row_live = row_ok > 0 ? (unsigned)Wr : 0u;
                int col = -slot_r;
#pragma unroll 2
                for (int d = 0; d < ndiag; d++, col++, pix_d += pix_step, z_d += z_step) {
                    step(pix_d, z_d, (unsigned)col < row_live);
                    __syncwarp();
                }
            } else if (iters == 1) {
                // several warps, one row per thread: a warp whose rows are all off the front only
                // meets the others at the barrier
                const unsigned row_live = row_ok > 0 ? (unsigned)Wr : 0u;
                const int wfirst = __shfl_sync(0xffffffffu, slot_r, 0);       // rows of this warp
                const int wlast = __shfl_sync(0xffffffffu, slot_r, 31);
                const int d_on = wfirst, d_off = wlast + Wr;                  // live for d in [d_on, d_off)
                int col = -slot_r;
                for (int d = 0; d < ndiag; d++, col++, pix_d += pix_step, z_d += z_step) {
                    if (d >= d_on && d < d_off) step(pix_d, z_d, (unsigned)col < row_live);
                    asm volatile("bar.sync 1, %0;" ::"r"(nwork) : "memory");  // worker warps only
                }
            } else {
                for (int d = 0; d < ndiag; d++) {
                    uint32_t pix = pix_d, za = z_d;
                    int col = d - slot_r;           // column of this thread's row `it` on diagonal d
#pragma unroll 1
                    for (int it = 0; it < iters; it++, pix += pix_row, za += z_row, col -= nslots) {
                        const bool active = row_ok > it && (unsigned)col < (unsigned)Wr;
                        if (!__any_sync(0xffffffffu, active)) continue;        // warp-uniform
                        step(pix, za, active);
                    }
                    pix_d += pix_step;
                    z_d += z_step;
                    if (nwork <= 32) __syncwarp();
                    else asm volatile("bar.sync 1, %0;" ::"r"(nwork) : "memory");
                }
            }
        }

        IFK_PROBE(6);
        if (b_next < p.B) {                  // the stripe goes on: ybuf back to zero for the next image
            __syncthreads();
            for (int i = tid * 4; i < p.YN; i += nthr * 4)
                *reinterpret_cast<float4 *>(ybuf + i) = make_float4(0.f, 0.f, 0.f, 0.f);
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
            for (int i = tid; i < Cg * HW; i += nthr) dst[i] = zbuf[i];
            __syncthreads();
        }
    }
    IFK_PROBE(7);
    if (p.bulk && tid == 0) bulk_store_wait_read();   // smem must outlive the last store's read
    IFK_PROBE(8);
}

// one launcher per vector width (defined in ifk_solve_v{1,2,4}.cu); returns 0, a cudaError_t,
// or IFK_ERR_UNSUPPORTED when (cc, nv) is not an instantiated variant
int launch_solve_vec1(int cc, int nv, const SolveParams &p, dim3 grid, int threads, size_t smem, cudaStream_t s);
int launch_solve_vec2(int cc, int nv, const SolveParams &p, dim3 grid, int threads, size_t smem, cudaStream_t s);
int launch_solve_vec4(int cc, int nv, const SolveParams &p, dim3 grid, int threads, size_t smem, cudaStream_t s);
bool solve_use_pdl();
int solve_variant_max_threads(int cc, int nv, int vec);

template <int CC, int NV, int VEC>
int launch_solve_variant(const SolveParams &p, dim3 grid, int threads, size_t smem, cudaStream_t s)
{
    auto kern = solve_smem_kernel<CC, NV, VEC>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL, see the kernel prologue
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = solve_use_pdl() ? 1 : 0;
    return cuda_status(cudaLaunchKernelEx(&cfg, kern, p));
}

}  // namespace ifk
