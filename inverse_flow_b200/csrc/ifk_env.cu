// IFK_* knobs: parsed once, cached (see ifk_env.cuh).
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ifk_env.cuh"
#include "ifk_internal.cuh"

namespace ifk {

static EnvKnobs g_env;
static std::once_flag g_env_once;
static std::mutex g_env_mutex;
static unsigned g_env_generation = 0;

static bool is_one(const char *name)
{
    const char *e = getenv(name);
    return e && e[0] == '1';
}
static bool is_zero(const char *name)
{
    const char *e = getenv(name);
    return e && e[0] == '0';
}

static void parse_env()
{
    EnvKnobs k{};
    k.solve_global = is_one("IFK_SOLVE_GLOBAL");
    k.solve_stream = is_one("IFK_SOLVE_STREAM");
    k.solve_window = is_one("IFK_SOLVE_WINDOW");
    k.shfl_off = is_zero("IFK_SOLVE_SHFL");
    k.wave_off = is_zero("IFK_SOLVE_WAVE");
    k.split_off = !is_one("IFK_SOLVE_SPLIT");     // opt-in: measured slower than the wave kernel
    k.nobulk = is_one("IFK_SOLVE_NOBULK");
    k.dw_quad_off = is_zero("IFK_DW_QUAD");
    k.pdl = !is_zero("IFK_PDL");
    if (const char *e = getenv("IFK_SHFL_NCT")) k.shfl_nct = atoi(e);
    k.conv_wide = -1;
    if (const char *e = getenv("IFK_CONV_WIDE")) k.conv_wide = e[0] == '1' ? 1 : 0;
    if (const char *e = getenv("IFK_SOLVE_CFG"))
        k.has_solve_cfg = *e && sscanf(e, "%d,%d,%d,%d,%d", &k.solve_cfg[0], &k.solve_cfg[1], &k.solve_cfg[2],
                                       &k.solve_cfg[3], &k.solve_cfg[4]) == 5;
    if (const char *e = getenv("IFK_STREAM_CFG")) sscanf(e, "%d,%d", &k.stream_cfg[0], &k.stream_cfg[1]);
    if (const char *e = getenv("IFK_WINDOW_CFG"))
        sscanf(e, "%d,%d,%d,%d", &k.window_cfg[0], &k.window_cfg[1], &k.window_cfg[2], &k.window_cfg[3]);
    if (const char *e = getenv("IFK_WAVE_CFG"))
        sscanf(e, "%d,%d,%d,%d", &k.wave_cfg[0], &k.wave_cfg[1], &k.wave_cfg[2], &k.wave_cfg[3]);
    if (const char *e = getenv("IFK_DW_CFG")) sscanf(e, "%d,%d", &k.dw_cfg[0], &k.dw_cfg[1]);
    if (const char *e = getenv("IFK_SPLIT_CFG")) sscanf(e, "%d,%d", &k.split_cfg[0], &k.split_cfg[1]);
    if (const char *e = getenv("IFK_PREP_CFG")) sscanf(e, "%d,%d", &k.prep_cfg[0], &k.prep_cfg[1]);
    std::lock_guard<std::mutex> lock(g_env_mutex);
    g_env = k;
    g_env_generation++;
}

const EnvKnobs &env()
{
    std::call_once(g_env_once, parse_env);
    return g_env;
}

void reload_env()
{
    std::call_once(g_env_once, parse_env);
    parse_env();
}

static int query_attr(cudaDeviceAttr attr, int fallback)
{
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess || v <= 0) {
        cudaGetLastError();      // no device (host-only introspection): the B200 figures
        return fallback;
    }
    return v;
}

unsigned env_generation()
{
    std::call_once(g_env_once, parse_env);
    std::lock_guard<std::mutex> lock(g_env_mutex);
    return g_env_generation;
}

int device_sm_count()
{
    static int v = query_attr(cudaDevAttrMultiProcessorCount, kNumSM);
    return v;
}

int device_max_smem_optin()
{
    static int v = query_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, kMaxSmemBytes);
    return v;
}

}  // namespace ifk
