// Internal declarations shared by the kernel translation units.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ifk.h"

namespace ifk {

constexpr int kNumSM = 148;               // B200: 2 dies x 74 SMs
constexpr int kMaxSmemBytes = 227 * 1024; // opt-in dynamic shared memory per CTA

struct Geometry {
    int B, C, H, W, KH, KW, Cw, groups;
    int orient;  // IFK_ORIENT_*: bit 0 reflects W, bit 1 reflects H
    int flags;   // IFK_FLAG_*
    int Cg;   // channels per group
    int K;    // KH * KW taps (tap 0 = the centre / "x" tap)
    int KD;   // K * Cg : length of one prepared weight row
    int KDP;  // KD rounded up to a multiple of 4 (row stride of the prepared weight)
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int make_geometry(const ifk_problem *p, Geometry *g);          // validates; IFK_ERR_* or 0
size_t prepared_floats(const Geometry &g);                      // canonical rows + the wave kernel's packed copy
int launch_wave_pack(const Geometry &g, float *prepared, int count, size_t prepared_stride, cudaStream_t s);
inline const float *prepared_dir(const Geometry &g, const float *prepared, int dir) {
    return prepared + (size_t)dir * g.C * g.KDP;
}

// launchers (each returns 0 or a cudaError_t)
size_t prepare_smem_bytes(int Cg);      // dynamic shared memory of prepare_kernel (ifk_prepare.cu)
int launch_prepare(const Geometry &g, const float *weight, float *prepared, cudaStream_t s, int count = 1,
                   size_t weight_stride = 0, size_t prepared_stride = 0);
int launch_solve(const Geometry &g, const float *in, const float *prepared, float *out,
                 bool reverse, cudaStream_t s, long long *probe = nullptr);     // `prepared`: the whole buffer
int launch_conv(const Geometry &g, const float *y, const float *weight, float *x, cudaStream_t s);
size_t bwd_weight_workspace_bytes(const Geometry &g);
int launch_bwd_weight(const Geometry &g, const float *dx, const float *y, float *dw,
                      void *workspace, cudaStream_t s);
int launch_bwd_weight_partial(const Geometry &g, const float *dx, const float *y, void *workspace, cudaStream_t s);
int launch_bwd_weight_reduce(const Geometry &g, int count, const void *workspace, size_t workspace_stride,
                             float *dw, size_t dw_stride, cudaStream_t s);
int describe_solve(const Geometry &g, char *buf, size_t buflen);
bool stream_solve_available(const Geometry &g);
int describe_stream_solve(const Geometry &g, char *buf, size_t buflen);
int launch_solve_stream(const Geometry &g, const float *in, const float *prep_dir, float *out, bool reverse,
                        cudaStream_t s);
bool window_solve_available(const Geometry &g);
int describe_window_solve(const Geometry &g, char *buf, size_t buflen);
int launch_solve_window(const Geometry &g, const float *in, const float *prep_dir, float *out, bool reverse,
                        cudaStream_t s);
bool shfl_solve_available(const Geometry &g);
int describe_shfl_solve(const Geometry &g, char *buf, size_t buflen);
int launch_solve_shfl(const Geometry &g, const float *in, const float *prep_dir, float *out, bool reverse,
                      long long *probe, cudaStream_t s);
int launch_solve_chain(const Geometry &g, int n, const int *orients, const float *const *prepared, const float *x,
                       float *const *ys, cudaStream_t s);
bool wave_solve_available(const Geometry &g);
int describe_wave_solve(const Geometry &g, char *buf, size_t buflen);
int launch_solve_wave(const Geometry &g, const float *in, const float *prepared, float *out, bool reverse,
                      int flags, long long *probe, cudaStream_t s);
// fused neighbours (ifk.h: ifk_fused); out may be nullptr when out2 is given
int launch_solve_wave_fused(const Geometry &g, const ifk_fused &f, const float *in, const float *prepared, float *out,
                            float *out2, bool reverse, cudaStream_t s);

bool split_solve_available(const Geometry &g);
int describe_split_solve(const Geometry &g, char *buf, size_t buflen);
size_t split_pack_floats(const Geometry &g);                    // per layer, floats
size_t split_pack_offset(const Geometry &g);                    // floats from the start of a layer's prepared buffer
int launch_split_pack(const Geometry &g, const float *prepared, float *pack, int count, size_t prepared_stride,
                      cudaStream_t s);
int launch_split_layers(const Geometry &g, int n, const int *orients, const float *const *packs, const float *in,
                        float *const *outs, bool reverse, int flags, long long *probe, cudaStream_t s);

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? 0 : (int)e; }

}  // namespace ifk
