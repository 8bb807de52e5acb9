// extern "C" boundary (include/ifk.h): argument validation and dispatch only.
#include <stdio.h>
#include "ifk_env.cuh"
#include "ifk_internal.cuh"

namespace ifk {

int make_geometry(const ifk_problem *p, Geometry *g)
{
    if (!p) return IFK_ERR_NULL_POINTER;
    if (p->B < 0 || p->C <= 0 || p->H <= 0 || p->W <= 0 || p->KH <= 0 || p->KW <= 0 || p->Cw <= 0)
        return IFK_ERR_BAD_SHAPE;
    if (p->groups < 1 || p->C % p->groups != 0 || p->Cw < p->C / p->groups) return IFK_ERR_BAD_GROUPS;
    g->B = p->B; g->C = p->C; g->H = p->H; g->W = p->W; g->KH = p->KH; g->KW = p->KW;
    if (p->orient < 0 || p->orient > 3) return IFK_ERR_BAD_ORIENT;
    if (p->flags & ~IFK_FLAG_STABLE_PREPARED) return IFK_ERR_BAD_FLAGS;
    g->Cw = p->Cw; g->groups = p->groups; g->orient = p->orient; g->flags = p->flags;
    g->Cg = p->C / p->groups;
    g->K = p->KH * p->KW;
    // 32-bit index arithmetic inside one image / one weight tensor
    const long long per_image = (long long)p->C * p->H * p->W;
    const long long wsize = (long long)p->C * p->Cw * g->K;
    const long long kd = (long long)g->K * g->Cg;
    if (per_image >= (1LL << 30) || wsize >= (1LL << 30) || kd >= (1LL << 24) || p->groups > 65535)
        return IFK_ERR_UNSUPPORTED;
    g->KD = (int)kd;
    g->KDP = round_up(g->KD, 4);
    if (prepare_smem_bytes(g->Cg) > (size_t)kMaxSmemBytes)
        return IFK_ERR_UNSUPPORTED;                      // ifk_prepare stages three Cg x Cg matrices
    return IFK_OK;
}

}  // namespace ifk

using namespace ifk;

extern "C" {

int ifk_version(void) { return IFK_VERSION; }

const char *ifk_status_string(int status)
{
    switch (status) {
        case IFK_OK: return "ok";
        case IFK_ERR_NULL_POINTER: return "a required pointer is NULL";
        case IFK_ERR_BAD_SHAPE: return "bad shape: negative dimension or zero C/H/W/KH/KW/Cw";
        case IFK_ERR_BAD_GROUPS: return "bad groups: need groups >= 1, C % groups == 0, Cw >= C/groups";
        case IFK_ERR_UNSUPPORTED: return "shape not supported by the kernels";
        case IFK_ERR_NO_DEVICE: return "no CUDA device";
        case IFK_ERR_BAD_ORIENT: return "bad orient: need one of IFK_ORIENT_TL/TR/BL/BR (0..3)";
        case IFK_ERR_BAD_LAYOUT: return "bad layout: tensors must be NCHW-contiguous float32 (channels_last is not accepted)";
        case IFK_ERR_BAD_FLAGS: return "bad flags: unknown IFK_FLAG_* bits";
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown ifk status";
}

size_t ifk_prepared_floats(const ifk_problem *p)
{
    Geometry g;
    if (make_geometry(p, &g) != IFK_OK) return 0;
    return prepared_floats(g);
}

int ifk_prepare_f32(const ifk_problem *p, const float *weight, float *prepared, ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!weight || !prepared) return IFK_ERR_NULL_POINTER;
    return launch_prepare(g, weight, prepared, (cudaStream_t)stream);
}

int ifk_prepare_many_f32(const ifk_problem *p, int count, const float *weights, size_t weight_stride,
                         float *prepared, size_t prepared_stride, ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (count < 0) return IFK_ERR_BAD_SHAPE;
    if (count > 0 && (!weights || !prepared)) return IFK_ERR_NULL_POINTER;
    if ((long long)count * g.groups > 0x7fffffffLL) return IFK_ERR_UNSUPPORTED;
    return launch_prepare(g, weights, prepared, (cudaStream_t)stream, count, weight_stride, prepared_stride);
}

int ifk_inverse_f32(const ifk_problem *p, const float *x, const float *prepared, float *y,
                    ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!prepared || (g.B > 0 && (!x || !y))) return IFK_ERR_NULL_POINTER;
    return launch_solve(g, x, prepared, y, false, (cudaStream_t)stream);
}

int ifk_conv_f32(const ifk_problem *p, const float *y, const float *weight, float *x,
                 ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!weight || (g.B > 0 && (!x || !y))) return IFK_ERR_NULL_POINTER;
    return launch_conv(g, y, weight, x, (cudaStream_t)stream);
}

int ifk_bwd_input_f32(const ifk_problem *p, const float *grad, const float *prepared, float *dx,
                      ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!prepared || (g.B > 0 && (!grad || !dx))) return IFK_ERR_NULL_POINTER;
    return launch_solve(g, grad, prepared, dx, true, (cudaStream_t)stream);
}

size_t ifk_bwd_weight_workspace_bytes(const ifk_problem *p)
{
    Geometry g;
    if (make_geometry(p, &g) != IFK_OK) return 0;
    return bwd_weight_workspace_bytes(g);
}

int ifk_bwd_weight_f32(const ifk_problem *p, const float *dx, const float *y, float *dw,
                       void *workspace, ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!dw || !workspace || (g.B > 0 && (!dx || !y))) return IFK_ERR_NULL_POINTER;
    return launch_bwd_weight(g, dx, y, dw, workspace, (cudaStream_t)stream);
}

int ifk_bwd_weight_partial_f32(const ifk_problem *p, const float *dx, const float *y, void *workspace,
                               ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!workspace || (g.B > 0 && (!dx || !y))) return IFK_ERR_NULL_POINTER;
    return launch_bwd_weight_partial(g, dx, y, workspace, (cudaStream_t)stream);
}

int ifk_bwd_weight_reduce_many_f32(const ifk_problem *p, int count, const void *workspaces,
                                   size_t workspace_stride_bytes, float *dw, size_t dw_stride,
                                   ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (count < 0 || count > 65535 || workspace_stride_bytes % sizeof(float) != 0) return IFK_ERR_BAD_SHAPE;
    if (count > 0 && (!workspaces || !dw)) return IFK_ERR_NULL_POINTER;
    return launch_bwd_weight_reduce(g, count, workspaces, workspace_stride_bytes, dw, dw_stride,
                                    (cudaStream_t)stream);
}

int ifk_backward_f32(const ifk_problem *p, const float *grad, const float *y, const float *prepared,
                     float *dx, float *dw, void *workspace, ifk_stream_t stream)
{
    int st = ifk_bwd_input_f32(p, grad, prepared, dx, stream);
    if (st != IFK_OK) return st;
    return ifk_bwd_weight_f32(p, dx, y, dw, workspace, stream);
}

int ifk_inverse_once_f32(const ifk_problem *p, const float *x, const float *weight, float *scratch, float *y,
                         ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!weight || !scratch || (g.B > 0 && (!x || !y))) return IFK_ERR_NULL_POINTER;
    st = launch_prepare(g, weight, scratch, (cudaStream_t)stream);
    if (st != IFK_OK) return st;
    g.flags &= ~IFK_FLAG_STABLE_PREPARED;        // the prepare kernel is this solve's predecessor
    return launch_solve(g, x, scratch, y, false, (cudaStream_t)stream);
}

int ifk_inverse_chain_f32(const ifk_problem *p, int n, const int *orients, const float *const *prepared,
                          const float *x, float *const *ys, ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (n < 0) return IFK_ERR_BAD_SHAPE;
    if (n == 0) return IFK_OK;
    if (!orients || !prepared || !ys || (g.B > 0 && !x)) return IFK_ERR_NULL_POINTER;
    for (int i = 0; i < n; i++) {
        if (orients[i] < 0 || orients[i] > 3) return IFK_ERR_BAD_ORIENT;
        if (!prepared[i] || (g.B > 0 && !ys[i])) return IFK_ERR_NULL_POINTER;
    }
    return launch_solve_chain(g, n, orients, prepared, x, ys, (cudaStream_t)stream);
}

static int check_fused(const Geometry &g, const ifk_fused *f)
{
    if (!f) return IFK_ERR_NULL_POINTER;
    if (f->squeeze != 0 && f->squeeze != 1) return IFK_ERR_BAD_FLAGS;
    if (f->squeeze && g.Cg % 4 != 0) return IFK_ERR_BAD_SHAPE;
    if (!wave_solve_available(g)) return IFK_ERR_UNSUPPORTED;
    return IFK_OK;
}

int ifk_inverse_fused_f32(const ifk_problem *p, const ifk_fused *f, const float *x, const float *prepared, float *y,
                          ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if ((st = check_fused(g, f)) != IFK_OK) return st;
    if (!prepared || (g.B > 0 && (!x || !y))) return IFK_ERR_NULL_POINTER;
    if (g.B == 0) return IFK_OK;
    return launch_solve_wave_fused(g, *f, x, prepared, y, nullptr, false, (cudaStream_t)stream);
}

int ifk_bwd_input_fused_f32(const ifk_problem *p, const ifk_fused *f, const float *grad, const float *prepared,
                            float *dx, float *dz, ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if ((st = check_fused(g, f)) != IFK_OK) return st;
    if (!prepared || (g.B > 0 && (!grad || !dz))) return IFK_ERR_NULL_POINTER;
    if (g.B == 0) return IFK_OK;
    return launch_solve_wave_fused(g, *f, grad, prepared, dx, dz, true, (cudaStream_t)stream);
}

// Measuring aid (ifk.h): one solve whose CTA (0,0) stamps clock64() at its phase boundaries.
int ifk_inverse_probe_f32(const ifk_problem *p, const float *x, const float *prepared, float *y,
                          long long *probe, ifk_stream_t stream)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!prepared || !probe || (g.B > 0 && (!x || !y))) return IFK_ERR_NULL_POINTER;
    return launch_solve(g, x, prepared, y, false, (cudaStream_t)stream, probe);
}

void ifk_debug_reload_env(void) { reload_env(); }

int ifk_describe_solve(const ifk_problem *p, char *buf, size_t buflen)
{
    Geometry g;
    int st = make_geometry(p, &g);
    if (st != IFK_OK) return st;
    if (!buf || buflen == 0) return IFK_ERR_NULL_POINTER;
    return describe_solve(g, buf, buflen);
}

}  // extern "C"
