// Test / tuning knobs (IFK_* environment variables), read ONCE per process and cached: the
// product path never calls getenv() on a launch.  ifk_debug_reload_env() (ifk.h) re-reads them --
// that is how the test-suite pins a particular kernel inside one process.
#pragma once

namespace ifk {

struct EnvKnobs {
    bool solve_global;   // IFK_SOLVE_GLOBAL=1  : the plain fallback solve kernel
    bool solve_stream;   // IFK_SOLVE_STREAM=1  : the stream kernel
    bool solve_window;   // IFK_SOLVE_WINDOW=1  : the window kernel
    bool shfl_off;       // IFK_SOLVE_SHFL=0    : no shuffle kernel
    bool wave_off;       // IFK_SOLVE_WAVE=0    : no pipelined wavefront kernel (-> the older resident kernel)
    bool split_off;      // IFK_SOLVE_SPLIT=1 switches the warp-specialised wavefront kernel ON (default: the wave kernel)
    bool nobulk;         // IFK_SOLVE_NOBULK=1  : no TMA bulk staging
    bool dw_quad_off;    // IFK_DW_QUAD=0       : dW stage 1 by the per-tap kernel also where the quad kernel applies
    bool pdl;            // IFK_PDL=0 switches programmatic dependent launch off
    int shfl_nct;        // IFK_SHFL_NCT        : tuning
    int conv_wide;       // IFK_CONV_WIDE       : -1 unset, 0 / 1 forced
    bool has_solve_cfg;  // IFK_SOLVE_CFG="cc,nv,vec,ns,nslots"
    int solve_cfg[5];
    int stream_cfg[2];   // IFK_STREAM_CFG="cc,nv"          (0 = unset)
    int window_cfg[4];   // IFK_WINDOW_CFG="cc,nv,cs,rp"    (0 = unset)
    int wave_cfg[4];     // IFK_WAVE_CFG="cc,ns,vec,threads" (0 = unset) : tuning
    int dw_cfg[2];       // IFK_DW_CFG="warps,ctas"       (0 = unset) : tuning of the quad dW kernel
    int split_cfg[2];    // IFK_SPLIT_CFG="nsh"            (0 = unset) : tuning
    int prep_cfg[2];     // IFK_PREP_CFG="0,taps"         : taps per CTA of the tap products (0 = unset; first field reserved) : tuning
    bool pins_other_solver() const { return solve_global || solve_stream || solve_window || has_solve_cfg; }
};

const EnvKnobs &env();
void reload_env();
unsigned env_generation();    // bumped by reload_env(): memoised kernel selections start over

int device_sm_count();        // cudaDevAttrMultiProcessorCount of the current device (148 on B200; cached)
int device_max_smem_optin();  // cudaDevAttrMaxSharedMemoryPerBlockOptin (227 KB on B200; cached)

}  // namespace ifk
