/*
 * ifk.h -- C ABI of the B200 inverse-convolution kernels ("ifk" = inverse-flow kernels).
 *
 * This is the drop-in boundary for the ONE hot path this repository replaces: the
 * `inv_conv_with_bp` CUDA extension of girish-lab/Inverse-Flow, i.e. the four entry points
 * exported at inf/utils/inv_conv_cuda/inv_conv_with_bp_general.cpp:115-120
 * (`inverse`, `forward`, `dy`, `dw`) and called from inf/layers/inv_conv.py:52,74,79,266,459.
 *
 * Conventions (all functions):
 *   - plain C, no torch / ATen types; every pointer is a DEVICE pointer to float32 data
 *     owned by the caller; nothing is allocated, freed or retained by the library;
 *   - tensors are NCHW-contiguous (the layout the reference passes, inv_conv.py:45-60); a
 *     channels_last (NHWC) activation must be made contiguous by the caller -- the Python
 *     binding rejects it with IFK_ERR_BAD_LAYOUT instead of reading it as NCHW;
 *     the weight is (C, Cw, KH, KW) contiguous with Cw >= C/groups -- the reference layers
 *     pass Cw == C and the kernels read the first C/groups input columns of each row
 *     (inv_conv_with_bp_kernel_general.cu:62);
 *   - work is enqueued on the caller's stream and the call returns immediately: no device
 *     synchronisation (the reference calls cudaDeviceSynchronize() per diagonal, .cu:124),
 *     graph-capturable, re-entrant; the only process-wide state is the read-once cache of
 *     the IFK_* test / tuning environment knobs (see ifk_debug_reload_env);
 *   - return value: 0 = ok, negative = IFK_ERR_* argument error (nothing was launched),
 *     positive = a cudaError_t raised by a launch.
 *
 * Math contract.  With Cg = C/groups, `base` the first channel of
 * c's group and w[c][kc][q] := W[c][kc][KH-1-qh][KW-1-qw] for a shift q = (qh, qw):
 *
 *   (L y)[b,c,p] = y[b,c,p] + sum_{(q,kc) != (0,c-base), kc < c-base when q == 0}
 *                                 w[c][kc][q] * y[b, base+kc, p-q]        (zero outside)
 *
 * i.e. a causal (top-left zero padded) masked convolution with an implicit unit diagonal
 * and a strictly-lower-triangular centre tap.  `groups == 1` is the semantics of the
 * reference's CPU solvers (inf/utils/solve_mc.py:88-114, fastflow_inverse/
 * solve_parallel_mc.pyx:100-124, cinc_kernel_level1.cu:57-69); `groups == 4` that of
 * cinc_kernel_level2.cu:59-72 and, for C == 4, of the shipped inv_conv_with_bp kernels.
 */
#ifndef IFK_H
#define IFK_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFK_VERSION 400 /* major*10000 + minor*100 + patch */

/* cudaStream_t without the CUDA headers (a driver-level CUstream handle). */
typedef struct CUstream_st *ifk_stream_t;

enum ifk_status {
    IFK_OK = 0,
    IFK_ERR_NULL_POINTER = -1,  /* a required pointer is NULL                            */
    IFK_ERR_BAD_SHAPE = -2,     /* a dimension is negative, or C/H/W/KH/KW is zero        */
    IFK_ERR_BAD_GROUPS = -3,    /* groups < 1, C % groups != 0 or Cw < C/groups          */
    IFK_ERR_UNSUPPORTED = -4,   /* shape exceeds what the kernels address */
    IFK_ERR_NO_DEVICE = -5,     /* no CUDA device / driver                               */
    IFK_ERR_BAD_ORIENT = -6,    /* orient is not one of IFK_ORIENT_*                     */
    IFK_ERR_BAD_LAYOUT = -7,    /* a tensor is not NCHW-contiguous float32 (e.g. channels_last) */
    IFK_ERR_BAD_FLAGS = -8      /* unknown bits in ifk_problem.flags                     */
};

/* ifk_problem.flags */
enum ifk_flags {
    /* The caller vouches that the operation enqueued on `stream` immediately before this call did
     * NOT write the `prepared` buffer the call reads (true for every solve of a layer chain except
     * the first one behind ifk_prepare*).  The solve kernels then fetch their weights ahead of the
     * programmatic-dependent-launch wait, overlapping the previous kernel; without the flag the
     * fetch happens after the wait (CUDA only guarantees a predecessor's writes after it). */
    IFK_FLAG_STABLE_PREPARED = 1
};

/* Corner the causal support grows from (the reference layers' `order`, inf/layers/inv_conv.py:
 * 198-214: they flip the image -- and the stored weight -- around a top-left kernel).  Here the
 * flips are index reflections inside the kernels, nothing is copied: with F the reflection of
 * the listed axes, every operator below becomes F (.) F, e.g. ifk_inverse_f32 computes
 * F L^-1 F x.  bit 0 = reflect W, bit 1 = reflect H. */
enum ifk_orient {
    IFK_ORIENT_TL = 0,  /* top-left (no reflection): inv_flow_no_pad and order 'TL' */
    IFK_ORIENT_TR = 1,  /* reflect W */
    IFK_ORIENT_BL = 2,  /* reflect H */
    IFK_ORIENT_BR = 3   /* reflect both */
};

/* Geometry of one call.  B may be 0 (the call is a no-op that still validates). */
typedef struct ifk_problem {
    int B, C, H, W; /* activations: (B, C, H, W)                                */
    int KH, KW;     /* kernel taps                                              */
    int Cw;         /* second dimension of the weight tensor, >= C / groups     */
    int groups;     /* channel groups; the reference kernels hard-code 4        */
    int orient;     /* IFK_ORIENT_*; 0 for the plain top-left operator          */
    int flags;      /* IFK_FLAG_* bits; 0 is always correct                     */
} ifk_problem;

int ifk_version(void);
const char *ifk_status_string(int status);

/* ---- prepared weights ---------------------------------------------------------------
 * The solve kernels consume a "prepared" copy of the weight: per group, the non-centre
 * taps pre-multiplied by T = (I + A0)^-1 (A0 = the strictly-lower centre tap) and negated,
 * for the forward solve (L^-1) and, transposed, for the adjoint solve (L^-T).  Folding T
 * removes the per-pixel Cg-step channel substitution from the wavefront's critical path.
 * One ifk_prepare_f32 call serves one ifk_inverse_f32 and the matching ifk_backward_f32.
 * `prepared` must hold ifk_prepared_floats(p) floats and should start at a multiple of 16 bytes
 * (the model-shape solve kernels read their packed copy as 16-byte words and refuse an unaligned
 * buffer with IFK_ERR_UNSUPPORTED).  Internally: T once per layer, the tap products, the packed
 * copy -- three launches on `stream`, chained by programmatic dependent launch. */
size_t ifk_prepared_floats(const ifk_problem *p);
int ifk_prepare_f32(const ifk_problem *p, const float *weight, float *prepared,
                    ifk_stream_t stream);

/* The same for `count` weight tensors of one geometry in ONE launch: tensor i starts at
 * weights + i*weight_stride and is prepared into prepared + i*prepared_stride (strides in
 * floats).  A model's layers do not depend on activations, so a training step prepares all
 * layers of a stage up front with one call. */
int ifk_prepare_many_f32(const ifk_problem *p, int count, const float *weights, size_t weight_stride,
                         float *prepared, size_t prepared_stride, ifk_stream_t stream);

/* y = L^-1 x : the layer's training direction x -> z.
 * Replaces inv_conv_with_bp.inverse (inv_conv_with_bp_general.cpp:19-28 ->
 * inv_conv_cuda_inverse, inv_conv_with_bp_kernel_general.cu:72-129).  x and y must not alias. */
int ifk_inverse_f32(const ifk_problem *p, const float *x, const float *prepared, float *y,
                    ifk_stream_t stream);

/* x = L y : the masked convolution, sampling direction z -> x.
 * Replaces inv_conv_with_bp.forward (inv_conv_with_bp_general.cpp:44-53 ->
 * inv_conv_fwd_cuda_inverse, .cu:203-264).  Takes the RAW weight.  x and y must not alias. */
int ifk_conv_f32(const ifk_problem *p, const float *y, const float *weight, float *x,
                 ifk_stream_t stream);

/* dX = L^-T g : gradient w.r.t. the layer input.
 * Replaces inv_conv_with_bp.dy (inv_conv_with_bp_general.cpp:70-81 -> inv_conv_dy,
 * .cu:388-483), computing the true adjoint.  g and dx must not alias. */
int ifk_bwd_input_f32(const ifk_problem *p, const float *g, const float *prepared, float *dx,
                      ifk_stream_t stream);

/* dW[c][kc][KH-1-qh][KW-1-qw] = - sum_{b,p} dX[b,c,p] * y[b,base+kc,p-q]; masked taps and
 * columns kc >= C/groups are written as 0.  `y` is the saved OUTPUT of ifk_inverse_f32.
 * Replaces inv_conv_with_bp.dw (inv_conv_with_bp_general.cpp:99-112 -> inv_conv_dw,
 * .cu:634-735).  Deterministic: per-CTA partial sums in `workspace`, reduced in a fixed
 * order.  `workspace` must hold ifk_bwd_weight_workspace_bytes(p) bytes (16-byte aligned). */
size_t ifk_bwd_weight_workspace_bytes(const ifk_problem *p);
int ifk_bwd_weight_f32(const ifk_problem *p, const float *dx, const float *y, float *dw,
                       void *workspace, ifk_stream_t stream);

/* The two stages of ifk_bwd_weight_f32 separately, so a training step can run stage 1 of every
 * layer on a side stream while the next layer's dX solve proceeds, and finish `count` layers of
 * one geometry with ONE stage-2 launch: layer i reads workspaces + i*workspace_stride_bytes and
 * writes dw + i*dw_stride (floats). */
int ifk_bwd_weight_partial_f32(const ifk_problem *p, const float *dx, const float *y, void *workspace,
                               ifk_stream_t stream);
int ifk_bwd_weight_reduce_many_f32(const ifk_problem *p, int count, const void *workspaces,
                                   size_t workspace_stride_bytes, float *dw, size_t dw_stride,
                                   ifk_stream_t stream);

/* dX and dW in one call (what inv_conv_.backward needs, inv_conv.py:62-81). */
int ifk_backward_f32(const ifk_problem *p, const float *g, const float *y,
                     const float *prepared, float *dx, float *dw, void *workspace,
                     ifk_stream_t stream);

/* ---- introspection (used by bench.py / tests; not needed by a binding) -----------------
 * Which kernel variant a solve of this geometry dispatches to, as a short static string,
 * e.g. "smem<cc=4,nv=6,vec=4> ns=4 nct=3 slots=16 iters=1 threads=192(192) ..." or "global ...". */
int ifk_describe_solve(const ifk_problem *p, char *buf, size_t buflen);

/* y = L^-1 x from the RAW weight in one call: prepare + solve, the exact shape of the reference's
 * inverse(input, kernel, output) (inv_conv_with_bp_general.cpp:19-28).  `scratch` must hold
 * ifk_prepared_floats(p) floats; it is left holding the prepared weights (usable by a following
 * ifk_backward_f32). */
int ifk_inverse_once_f32(const ifk_problem *p, const float *x, const float *weight, float *scratch,
                         float *y, ifk_stream_t stream);

/* Consecutive inverse-conv layers that feed each other directly -- the TL/TR/BL/BR layers of
 * Inv_FlowUnit (inf/layers/inv_flow.py:28-53) -- as ONE launch: the image stays in shared memory
 * from layer to layer, every layer's output ys[i] is still written (the backward needs it).
 * Layer i uses orients[i] and prepared[i]; results are bit-identical to n ifk_inverse_f32 calls.
 * p->orient is ignored.  IFK_ERR_UNSUPPORTED when the geometry has no resident kernel
 * (callers then issue the n calls). */
int ifk_inverse_chain_f32(const ifk_problem *p, int n, const int *orients, const float *const *prepared,
                          const float *x, float *const *ys, ifk_stream_t stream);

/* ---- elementwise neighbours fused into a solve's load and store ------------
 * In the if_* Glow models every inverse-conv layer is preceded by an ActNorm affine and every block is
 * opened by a Squeeze (model order: inf/experiments/if_glow_mnist.py:62-124; ActNorm.forward
 * `(x - translation) * exp(-log_scale)`, inf/layers/actnorm.py:36; space_to_depth, inf/layers/squeeze.py:5-13).
 * Both are pure re-indexing / per-channel FMAs: the solve kernels apply them while the image moves between
 * global and shared memory, which removes two elementwise kernels (and their HBM round trips) per layer.
 *   ifk_inverse_fused_f32   : y  = L^-1( in_scale (.) S(x) + in_bias )
 *   ifk_bwd_input_fused_f32 : dx = L^-T g  (raw, what ifk_bwd_weight_f32 needs; may be NULL)
 *                             dz = S^T( out_scale (.) dx )   (the gradient handed to the layer before)
 * with S = space_to_depth when `squeeze` is set (x and dz are then (B, C/4, 2H, 2W) tensors; needs
 * (C/groups) % 4 == 0) and the identity otherwise.  Scale / bias vectors have C floats; NULL = 1 / 0.
 * IFK_ERR_UNSUPPORTED when the geometry is not served by the pipelined wavefront kernel (callers then run
 * the unfused sequence). */
typedef struct ifk_fused {
    const float *in_scale, *in_bias; /* forward: per-channel affine of the (squeezed) input        */
    const float *out_scale;          /* backward: per-channel scale of dz                           */
    int squeeze;                     /* 0 / 1                                                       */
} ifk_fused;
int ifk_inverse_fused_f32(const ifk_problem *p, const ifk_fused *f, const float *x, const float *prepared,
                          float *y, ifk_stream_t stream);
int ifk_bwd_input_fused_f32(const ifk_problem *p, const ifk_fused *f, const float *g, const float *prepared,
                            float *dx, float *dz, ifk_stream_t stream);

/* Phase timing of ONE solve (a debugging / measuring entry point, not part of the product path):
 * like ifk_inverse_f32, and thread 0 of CTA (0,0) writes clock64() stamps into `probe` (16 x int64
 * of device memory) -- 0 kernel start, 1..3 prologue done, 4 image landed, 5 / 6 diagonal loop
 * start / end, 7 store issued, 8 end. */
int ifk_inverse_probe_f32(const ifk_problem *p, const float *x, const float *prepared, float *y,
                          long long *probe, ifk_stream_t stream);

/* ---- data-parallel gradient exchange ---------------------------------------------
 * out[i] = sum over ranks r = 0..world-1 of buckets[r][i], summed in rank order on every rank (bit-identical
 * results everywhere), as ONE kernel over peer memory: it hand-shakes with the other ranks, reads their
 * buckets through NVLink/NVSwitch peer mappings and hand-shakes again before anyone may overwrite a bucket.
 * Replaces the gradient reduction of the reference's nn.DataParallel (inf/if_multiGPU_imagenet32.py:410-411).
 *   buckets[r] : rank r's bucket (n floats, 16-byte aligned) as mapped into THIS process (buckets[rank] = own)
 *   flags[r]   : rank r's flag block, ifk_allreduce_flag_bytes() bytes, zeroed once before the first call
 *   out        : local result, n floats, 16-byte aligned, not aliasing any bucket
 * Every rank must call it with the same n, the same number of times and in the same order; one process per
 * GPU (kernels of different ranks must be able to run at the same time).  Graph-capturable. */
size_t ifk_allreduce_flag_bytes(void);
int ifk_allreduce_peer_f32(const float *const *buckets, unsigned *const *flags, int rank, int world, float *out,
                           size_t n, ifk_stream_t stream);

/* Measured denominators for the roofline (bench.py, tools/hw_microbench.py).  Both calls synchronise the
 * device and are measuring aids, not part of the product path.
 *   ifk_debug_fp32_peak: TFLOP/s of independent FP32 FMAs on all SMs -- out_tflops[0] scalar FFMA,
 *     out_tflops[1] packed FFMA2; `scratch` >= 4 bytes of device memory.
 *   ifk_debug_latencies: cycles per DEPENDENT operation, one CTA: host_out[0] FFMA, [1] FFMA2, [2] warp
 *     shuffle, [3] shared-memory load, [4] st.shared -> __syncwarp -> ld.shared, [5] st.shared ->
 *     bar.sync (8 warps) -> ld.shared, [6] bar.sync alone (8 warps), [7] FADD; `device_out` 16 x int64. */
int ifk_debug_fp32_peak(float *scratch, double *out_tflops);
int ifk_debug_latencies(long long *device_out, double *host_out);

/* Re-read the IFK_* environment knobs (kernel pinning for tests and tuning runs).  They are parsed
 * once per process and cached: no launch ever calls getenv(). */
void ifk_debug_reload_env(void);

#ifdef __cplusplus
}
#endif
#endif /* IFK_H */
