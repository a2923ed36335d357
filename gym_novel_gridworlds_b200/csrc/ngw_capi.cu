// ngw_capi.cu — handle management and the extern "C" entry points of libngw_b200.so (include/ngw.h is the contract).
// The kernels live in ngw_step.cuh (hot path) and ngw_reset.cuh (cold paths); per-env device code in ngw_device.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ngw_reset.cuh"


// ====================================================================== host side: handle + C-ABI
using namespace ngw;

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));          \
    } while (0)

#define HOST_STREAMS 1

struct ngw_handle {
    int device = 0;
    long long n = 0, np = 0, first_gid = 0;
    unsigned long long seed = 0;
    int ms = 0, cells = 0, inv_stride = 0, obs_dim = 0, n_cfgs = 0;
    int map_bytes = 0, inv_bytes = 0, obs_bytes = 0, region_bytes = 0, warps = 4;
    bool use_tma = true, collect_stats = true, force_global_cfg = false, plain_store = false, use_pdl = true;
    bool pdl_in_graph = true;
    bool lidar_uniform = false;
    DevConfig* d_cfgs = nullptr;
    std::vector<int16_t*> d_luts;
    std::vector<DevConfig> h_cfgs;
    int8_t* map = nullptr; uchar4* pose = nullptr; int32_t* inv = nullptr; uint8_t* cfg_id = nullptr;
    uint32_t* episode = nullptr; int32_t* ep_len = nullptr; uint32_t* err = nullptr; double* stats = nullptr;
    uint8_t* zero_byte = nullptr;
    uint16_t* msg = nullptr;                       // caller-owned message-code buffer (ngw_set_message_buffer)
    int32_t* reset_list = nullptr; int32_t* reset_ctl = nullptr;   // auto-reset queue; ctl[0] = count, ctl[1] = finished CTAs
    int sm_count = 148;
    long long launches = 0;
    // host-buffer path
    cudaStream_t hs[HOST_STREAMS] = {nullptr};
    int32_t* h_actions = nullptr; int32_t* h_obs = nullptr; float* h_reward = nullptr; uint8_t* h_done = nullptr;
    float* h_cost = nullptr; uint8_t* h_result = nullptr;
};


extern "C" {

const char* ngw_last_error(void) { return g_err.c_str(); }
int ngw_abi_version(void) { return NGW_ABI_VERSION; }

void ngw_destroy(ngw_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto p : h->d_luts) cudaFree(p);
    cudaFree(h->d_cfgs); cudaFree(h->map); cudaFree(h->pose); cudaFree(h->inv); cudaFree(h->cfg_id);
    cudaFree(h->episode); cudaFree(h->ep_len); cudaFree(h->err); cudaFree(h->stats); cudaFree(h->zero_byte); cudaFree(h->reset_list); cudaFree(h->reset_ctl);
    cudaFree(h->h_actions); cudaFree(h->h_obs); cudaFree(h->h_reward);   // h_cost / h_done / h_result live in h_reward's block
    for (int i = 0; i < HOST_STREAMS; i++)
        if (h->hs[i]) cudaStreamDestroy(h->hs[i]);
    delete h;
}

static int create_init(ngw_handle* h, const ngw_config* cfgs, int32_t n_cfgs, int64_t n_envs, int32_t map_size,
                       int32_t device, int64_t first_env_gid, uint64_t seed, const cudaDeviceProp& prop);

int ngw_create(ngw_handle** out, const ngw_config* cfgs, int32_t n_cfgs, int64_t n_envs, int32_t map_size,
               int32_t device, int64_t first_env_gid, uint64_t seed) {
    if (!out || !cfgs || n_cfgs < 1 || n_cfgs > 255) return fail("ngw_create: need 1..255 configs");
    if (n_envs < 1) return fail("ngw_create: n_envs must be >= 1");
    if (map_size < 5 || map_size > NGW_MAX_MAP_SIZE) return fail("ngw_create: map_size out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("ngw_create: this library is built for sm_100a (B200) only");
    ngw_handle* h = new ngw_handle();
    h->device = device;
    if (create_init(h, cfgs, n_cfgs, n_envs, map_size, device, first_env_gid, seed, prop)) {
        std::string why = g_err;                    // ngw_destroy must not clobber the reason
        ngw_destroy(h);
        g_err = why;
        return 1;
    }
    *out = h;
    return 0;
}

static int create_init(ngw_handle* h, const ngw_config* cfgs, int32_t n_cfgs, int64_t n_envs, int32_t map_size,
                       int32_t device, int64_t first_env_gid, uint64_t seed, const cudaDeviceProp& prop) {
    h->device = device; h->n = n_envs; h->np = (n_envs + 31) / 32 * 32; h->first_gid = first_env_gid; h->seed = seed;
    h->ms = map_size; h->cells = map_size * map_size; h->n_cfgs = n_cfgs;
    h->use_tma = getenv("NGW_NO_TMA") == nullptr;
    h->collect_stats = getenv("NGW_NO_STATS") == nullptr;
    h->force_global_cfg = getenv("NGW_GLOBAL_CFG") != nullptr;
    h->plain_store = getenv("NGW_PLAIN_STORE") != nullptr;
    h->use_pdl = getenv("NGW_NO_PDL") == nullptr;
    h->pdl_in_graph = getenv("NGW_NO_PDL_GRAPH") == nullptr;
    for (int i = 0; i < n_cfgs; i++) {
        const ngw_config& c = cfgs[i];
        if (c.n_items < 1 || c.n_items > NGW_MAX_ITEMS || c.n_actions < 0 || c.n_actions > NGW_MAX_ACTIONS ||
            c.n_recipes > NGW_MAX_RECIPES || c.n_place > NGW_MAX_PLACE || c.n_reset_ops > NGW_MAX_RESET_OPS ||
            c.n_beams < 0 || c.max_range < 0 || c.n_beams * c.max_range > 4096) {
            return fail("ngw_create: config " + std::to_string(i) + " out of range");
        }
        if (c.n_items > h->inv_stride) h->inv_stride = c.n_items;
        int d = c.n_beams > 0 ? c.n_lidar_items * c.n_beams + c.n_inv_obs : 0;
        if (d > h->obs_dim) h->obs_dim = d;
        if (c.n_beams > 0 && c.beam_lut == nullptr) return fail("ngw_create: lidar config without beam_lut");
    }
    // device configs: the host beam LUT (d_row, d_col) becomes either the factorised 8-beam tables or an int16
    // linear-offset LUT for this map size
    h->h_cfgs.resize(n_cfgs);
    for (int i = 0; i < n_cfgs; i++) {
        DevConfig& dc = h->h_cfgs[i];
        memset(&dc, 0, sizeof(dc));
        dc.c = cfgs[i];
        const ngw_config& c = cfgs[i];
        const int B = c.n_beams, K = c.max_range;
        auto at = [&](int f, int b, int k, int j) { return (int)c.beam_lut[((f * B + b) * K + k) * 2 + j]; };
        bool fast = (B == 8 && K >= 1 && K <= NGW_MAX_RANGE);
        if (fast) {
            for (int par = 0; par < 2 && fast; par++)
                for (int k = 0; k < K && fast; k++) {
                    int dr = abs(at(0, par, k, 0)), dcol = abs(at(0, par, k, 1));
                    int d = dr > dcol ? dr : dcol;
                    if (d > 255) fast = false;
                    dc.lidar.disp[par][k] = (uint8_t)d;
                }
            for (int f = 0; f < 4 && fast; f++)
                for (int b = 0; b < 8 && fast; b++) {
                    int ur = at(f, b, 0, 0), uc = at(f, b, 0, 1);
                    if (abs(ur) > 1 || abs(uc) > 1 || (ur == 0 && uc == 0)) { fast = false; break; }
                    dc.lidar.unit[f][b] = (int16_t)(ur * map_size + uc);
                    for (int k = 0; k < K; k++) {
                        int d = dc.lidar.disp[b & 1][k];
                        if (at(f, b, k, 0) != ur * d || at(f, b, k, 1) != uc * d) { fast = false; break; }
                    }
                }
        }
        if (getenv("NGW_NO_FAST_LIDAR")) fast = false;
        dc.lidar.fast = fast ? 1 : 0;
        int16_t* d_lut = nullptr;
        if (B > 0 && !fast) {
            int n = 4 * B * K;
            std::vector<int16_t> lin(n);
            for (int j = 0; j < n; j++) lin[j] = (int16_t)(c.beam_lut[2 * j] * map_size + c.beam_lut[2 * j + 1]);
            CK(cudaMalloc(&d_lut, n * sizeof(int16_t)));
            CK(cudaMemcpy(d_lut, lin.data(), n * sizeof(int16_t), cudaMemcpyHostToDevice));
            h->d_luts.push_back(d_lut);
        }
        dc.lidar.lut = d_lut;
        dc.c.beam_lut = nullptr;
    }
    h->lidar_uniform = n_cfgs > 1;
    for (int i = 1; i < n_cfgs; i++) {
        const LidarDev &a = h->h_cfgs[0].lidar, &b = h->h_cfgs[i].lidar;
        if (!a.fast || !b.fast || h->h_cfgs[0].c.max_range != h->h_cfgs[i].c.max_range ||
            memcmp(a.unit, b.unit, sizeof(a.unit)) != 0 || memcmp(a.disp, b.disp, sizeof(a.disp)) != 0)
            h->lidar_uniform = false;
    }
    CK(cudaMalloc(&h->d_cfgs, sizeof(DevConfig) * n_cfgs));
    CK(cudaMemcpy(h->d_cfgs, h->h_cfgs.data(), sizeof(DevConfig) * n_cfgs, cudaMemcpyHostToDevice));
    // state
    CK(cudaMalloc(&h->map, (size_t)h->np * h->cells));
    CK(cudaMalloc(&h->pose, (size_t)h->np * 4));
    CK(cudaMalloc(&h->inv, (size_t)h->np * h->inv_stride * 4));
    CK(cudaMalloc(&h->cfg_id, (size_t)h->np));
    CK(cudaMalloc(&h->episode, (size_t)h->np * 4));
    CK(cudaMalloc(&h->ep_len, (size_t)h->np * 4));
    CK(cudaMalloc(&h->err, (size_t)h->np * 4));
    CK(cudaMalloc(&h->stats, sizeof(double) * NGW_STAT_SLOTS * NGW_STAT_COUNT));
    CK(cudaMalloc(&h->reset_list, (size_t)h->np * 4));
    CK(cudaMalloc(&h->reset_ctl, 16));
    CK(cudaMemset(h->reset_ctl, 0, 16));
    h->sm_count = prop.multiProcessorCount;
    CK(cudaMalloc(&h->zero_byte, 16));
    CK(cudaMemset(h->zero_byte, 0, 16));
    CK(cudaMemset(h->map, 0, (size_t)h->np * h->cells));
    CK(cudaMemset(h->pose, 0, (size_t)h->np * 4));
    CK(cudaMemset(h->inv, 0, (size_t)h->np * h->inv_stride * 4));
    CK(cudaMemset(h->cfg_id, 0, (size_t)h->np));
    CK(cudaMemset(h->episode, 0, (size_t)h->np * 4));
    CK(cudaMemset(h->ep_len, 0, (size_t)h->np * 4));
    CK(cudaMemset(h->err, 0, (size_t)h->np * 4));
    CK(cudaMemset(h->stats, 0, sizeof(double) * NGW_STAT_SLOTS * NGW_STAT_COUNT));
    // shared-memory carve-up per warp
    h->map_bytes = 32 * h->cells;                       // multiple of 32
    h->inv_bytes = 128 * h->inv_stride;
    h->obs_bytes = 128 * (h->obs_dim > 0 ? h->obs_dim : 0);
    h->region_bytes = 16 + 128 + 1024 + h->map_bytes + h->inv_bytes + h->obs_bytes;
    h->region_bytes = (h->region_bytes + 127) & ~127;
    if (h->region_bytes > 227 * 1024) return fail("ngw_create: map too large for shared memory");
    // G warps share one tile (warp 0 steps, all G cast 8/G lidar beams): 2 for small grids — the one-step kernel needs
    // 48 registers, so two-warp tiles of a 65,536-env batch are all resident — more when shared memory limits the tiles
    int tiles_per_sm = (227 * 1024) / (h->region_bytes + 1024);
    int warps = tiles_per_sm >= 6 ? 2 : (tiles_per_sm >= 3 ? 4 : 8);
    if (const char* w = getenv("NGW_WARPS")) {                      // tuning knob: warps per tile, 1 / 2 / 4 / 8
        int v = atoi(w);
        if (v == 1 || v == 2 || v == 3 || v == 4 || v == 8) warps = v;
    }
    h->warps = warps;
    CK(cudaFuncSetAttribute(step_kernel<true, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<true, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<true, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<true, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(step_kernel<false, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return 0;
}

int ngw_state(ngw_handle* h, ngw_state_view* o) {
    if (!h || !o) return fail("ngw_state: null");
    o->map = h->map; o->pose = reinterpret_cast<uint8_t*>(h->pose); o->inventory = h->inv; o->cfg_id = h->cfg_id;
    o->episode = h->episode; o->ep_len = h->ep_len; o->error_flags = h->err;
    o->inv_stride = h->inv_stride; o->obs_dim = h->obs_dim; o->n_envs = h->n; o->n_envs_padded = h->np;
    o->map_size = h->ms; o->n_configs = h->n_cfgs;
    return 0;
}

int ngw_set_env_configs(ngw_handle* h, const int32_t* cfg_id_dev, void* stream) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    int blocks = (int)((h->n + 255) / 256);
    set_cfg_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cfg_id_dev, h->cfg_id, h->n, h->n_cfgs, h->err);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int ngw_load_state(ngw_handle* h, const int8_t* map, const uint8_t* pose, const int32_t* inventory, int64_t first,
                   int64_t count, void* stream) {
    if (!h) return fail("null handle");
    if (first < 0 || count < 0 || first + count > h->n) return fail("ngw_load_state: range outside the batch");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (map) CK(cudaMemcpyAsync(h->map + first * h->cells, map, (size_t)count * h->cells, cudaMemcpyDeviceToDevice, s));
    if (pose) CK(cudaMemcpyAsync(h->pose + first, pose, (size_t)count * 4, cudaMemcpyDeviceToDevice, s));
    if (inventory)
        CK(cudaMemcpyAsync(h->inv + first * h->inv_stride, inventory, (size_t)count * h->inv_stride * 4,
                           cudaMemcpyDeviceToDevice, s));
    CK(cudaMemsetAsync(h->ep_len + first, 0, (size_t)count * 4, s));
    return 0;
}

static ResetParams reset_params(ngw_handle* h, const uint8_t* mask, int phase) {
    ResetParams p;
    p.dcfgs = h->d_cfgs; p.map = h->map; p.pose = h->pose; p.inv = h->inv; p.cfg_id = h->cfg_id; p.episode = h->episode;
    p.ep_len = h->ep_len; p.err = h->err; p.mask = mask; p.n_envs = h->n; p.first_gid = h->first_gid; p.seed = h->seed;
    p.ms = h->ms; p.cells = h->cells; p.inv_stride = h->inv_stride; p.phase = phase; p.zero_byte = h->zero_byte;
    p.reset_list = h->reset_list; p.reset_count = h->reset_ctl; p.done_ctas = h->reset_ctl + 1; p.obs = nullptr;
    p.obs_dim = h->obs_dim;
    return p;
}

int ngw_reset(ngw_handle* h, const uint8_t* mask, int32_t* obs, void* stream) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int blocks = (int)((h->n + 127) / 128);
    int rblocks = (int)((h->n + NGW_RESET_WARPS - 1) / NGW_RESET_WARPS);
    bool split = false;
    for (auto& c : h->h_cfgs) split |= c.c.reset_obs_after_ops < c.c.n_reset_ops;
    if (obs == nullptr || h->obs_dim == 0 || !split) {
        reset_kernel<<<rblocks, 32 * NGW_RESET_WARPS, 0, s>>>(reset_params(h, mask, 2));
        h->launches++;
        if (obs != nullptr && h->obs_dim > 0) {
            observe_masked_kernel<<<blocks, 128, 0, s>>>(reset_params(h, mask, 2), obs, h->obs_dim);
            h->launches++;
        }
    } else {
        reset_kernel<<<rblocks, 32 * NGW_RESET_WARPS, 0, s>>>(reset_params(h, mask, 0));
        observe_masked_kernel<<<blocks, 128, 0, s>>>(reset_params(h, mask, 0), obs, h->obs_dim);
        reset_kernel<<<rblocks, 32 * NGW_RESET_WARPS, 0, s>>>(reset_params(h, mask, 1));
        h->launches += 3;
    }
    CK(cudaGetLastError());
    return 0;
}

static StepParams step_params(ngw_handle* h, const int32_t* actions, int32_t* obs, float* reward, uint8_t* done,
                              float* cost, uint8_t* result, int auto_reset, int max_episode_steps, long long begin,
                              long long end) {
    StepParams p;
    p.dcfgs = h->d_cfgs; p.map = h->map; p.pose = h->pose; p.inv = h->inv; p.cfg_id = h->cfg_id; p.episode = h->episode;
    p.ep_len = h->ep_len; p.err = h->err; p.actions = actions; p.obs = h->obs_dim > 0 ? obs : nullptr; p.reward = reward;
    p.done = done; p.cost = cost; p.result = result; p.stats = h->collect_stats ? h->stats : nullptr;
    p.env_begin = begin; p.env_end = end; p.first_gid = h->first_gid; p.seed = h->seed;
    p.ms = h->ms; p.cells = h->cells; p.inv_stride = h->inv_stride; p.obs_dim = h->obs_dim;
    p.map_bytes = h->map_bytes; p.inv_bytes = h->inv_bytes; p.obs_bytes = h->obs_bytes;
    p.region_bytes = h->region_bytes; p.auto_reset = auto_reset; p.max_episode_steps = max_episode_steps;
    p.plain_store = h->plain_store ? 1 : 0;
    p.lidar_uniform = h->lidar_uniform ? 1 : 0;
    // streaming data (each tile is read once and its 8 KB of observations written once per step) should not linger in L2:
    // measured on C2 9.15 -> 8.82 us/step, C3 29.3 -> 28.0, C5 278 -> 274 (hinting the inventory store as well: 9.0)
    p.cache_hints = getenv("NGW_HINTS") ? atoi(getenv("NGW_HINTS")) : 3;
    p.msg = h->msg; p.reset_list = h->reset_list; p.reset_count = h->reset_ctl;
    p.policy_w = nullptr; p.policy_b = nullptr; p.policy_actions = 0;
    p.n_steps = 1; p.random_policy = 0; p.act_stride = 0; p.policy_seed = 0; p.done_count = nullptr; p.actions_out = nullptr;
    return p;
}

}  // extern "C" (templates need C++ linkage)

template <int NC>
static void launch_step_nc(ngw_handle* h, const StepParams& p, int blocks, size_t smem, cudaStream_t s) {
    static thread_local StepArgs<NC> args;          // host staging of the argument block (copied by the launch); per thread,
                                                    // so distinct handles stay independent across host threads
    args.p = p;
    for (int i = 0; i < NC && i < h->n_cfgs; i++) args.cfg[i] = h->h_cfgs[i];
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    const bool multi = p.n_steps > 1 || p.random_policy || p.done_count != nullptr || p.actions_out != nullptr ||
                       p.policy_w != nullptr;
    // the K-step rollout is all step logic (one lidar pass at the end): one warp per tile keeps more tiles resident
    const int warps = (multi && h->region_bytes * 12 <= 227 * 1024) ? 1 : h->warps;
    lc.gridDim = dim3(blocks); lc.blockDim = dim3(32 * warps); lc.dynamicSmemBytes = multi ? smem : smem - 1024;
    lc.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    // PDL (trigger after the tile's compute, see the kernel): eager python loop 12.3 -> 10.2 us/step, CUDA-graph replay
    // 9.95 -> 9.73 us/step on C2.  (An early trigger at kernel entry measured slower inside graphs.)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    lc.attrs = attr; lc.numAttrs = (h->use_pdl && (cap == cudaStreamCaptureStatusNone || h->pdl_in_graph)) ? 1 : 0;
    if (h->use_tma) {
        if (multi) cudaLaunchKernelEx(&lc, step_kernel<true, NC, true>, args);
        else cudaLaunchKernelEx(&lc, step_kernel<true, NC, false>, args);
    } else {
        if (multi) cudaLaunchKernelEx(&lc, step_kernel<false, NC, true>, args);
        else cudaLaunchKernelEx(&lc, step_kernel<false, NC, false>, args);
    }
}

static int launch_step(ngw_handle* h, const StepParams& p, cudaStream_t s) {
    long long tiles = (p.env_end - p.env_begin + 31) / 32;
    if (tiles <= 0) return 0;
    int blocks = (int)tiles;
    size_t smem = (size_t)h->region_bytes;
    int nc = h->force_global_cfg ? 0 : h->n_cfgs;
    if (nc == 0 || nc > 16) launch_step_nc<0>(h, p, blocks, smem, s);
    else if (nc == 1) launch_step_nc<1>(h, p, blocks, smem, s);
    else if (nc <= 4) launch_step_nc<4>(h, p, blocks, smem, s);
    else launch_step_nc<16>(h, p, blocks, smem, s);
    h->launches++;
    CK(cudaGetLastError());
    const bool multi = p.n_steps > 1 || p.random_policy || p.done_count != nullptr || p.actions_out != nullptr ||
                       p.policy_w != nullptr;
    if (p.actions != nullptr && p.auto_reset && !multi) {           // regenerate the episodes the step kernel queued
        ResetParams rp = reset_params(h, nullptr, 2);
        rp.obs = p.obs;
        reset_list_kernel<<<h->sm_count * 4, 32 * NGW_RESET_WARPS, 0, s>>>(rp);
        h->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" {

int ngw_step(ngw_handle* h, const int32_t* actions, int32_t* obs, float* reward, uint8_t* done, float* step_cost,
             uint8_t* result, int32_t auto_reset, int32_t max_episode_steps, void* stream) {
    if (!h) return fail("null handle");
    if (!actions || !reward || !done || !step_cost || !result) return fail("ngw_step: null output/action pointer");
    if (h->obs_dim > 0 && obs && ((uintptr_t)obs & 15)) return fail("ngw_step: obs must be 16-byte aligned");
    CK(cudaSetDevice(h->device));
    return launch_step(h, step_params(h, actions, obs, reward, done, step_cost, result, auto_reset, max_episode_steps,
                                      0, h->n), (cudaStream_t)stream);
}

int ngw_rollout(ngw_handle* h, const int32_t* actions, int32_t n_steps, uint64_t policy_seed, int32_t* obs,
                float* reward_sum, float* cost_sum, int32_t* done_count, uint8_t* last_done, uint8_t* last_result,
                int32_t* actions_out, int32_t auto_reset, int32_t max_episode_steps, void* stream) {
    if (!h) return fail("null handle");
    if (n_steps < 1) return fail("ngw_rollout: n_steps must be >= 1");
    if (!reward_sum || !cost_sum || !last_done || !last_result) return fail("ngw_rollout: null output pointer");
    if (h->obs_dim > 0 && obs && ((uintptr_t)obs & 15)) return fail("ngw_rollout: obs must be 16-byte aligned");
    CK(cudaSetDevice(h->device));
    // any non-null pointer marks "stepping"; with the random policy it is never dereferenced
    const int32_t* act = actions ? actions : reinterpret_cast<const int32_t*>(h->zero_byte);
    StepParams p = step_params(h, act, obs, reward_sum, last_done, cost_sum, last_result, auto_reset, max_episode_steps,
                               0, h->n);
    p.n_steps = n_steps; p.random_policy = actions ? 0 : 1; p.act_stride = h->n; p.policy_seed = policy_seed;
    p.done_count = done_count; p.actions_out = actions_out;
    return launch_step(h, p, (cudaStream_t)stream);
}

int ngw_rollout_policy(ngw_handle* h, const int32_t* weights, const int32_t* bias, int32_t n_policy_actions,
                       int32_t n_steps, int32_t* obs, float* reward_sum, float* cost_sum, int32_t* done_count,
                       uint8_t* last_done, uint8_t* last_result, int32_t* actions_out, int32_t auto_reset,
                       int32_t max_episode_steps, void* stream) {
    if (!h) return fail("null handle");
    if (!weights || !bias || n_policy_actions < 1 || n_policy_actions > 16)
        return fail("ngw_rollout_policy: need weights, bias and 1..16 policy actions");
    if (h->obs_dim == 0 || !obs) return fail("ngw_rollout_policy: needs a LidarInFront observation (and an obs buffer)");
    if (n_steps < 1) return fail("ngw_rollout_policy: n_steps must be >= 1");
    if (!reward_sum || !cost_sum || !last_done || !last_result) return fail("ngw_rollout_policy: null output pointer");
    if ((uintptr_t)obs & 15) return fail("ngw_rollout_policy: obs must be 16-byte aligned");
    CK(cudaSetDevice(h->device));
    StepParams p = step_params(h, reinterpret_cast<const int32_t*>(h->zero_byte), obs, reward_sum, last_done, cost_sum,
                               last_result, auto_reset, max_episode_steps, 0, h->n);
    p.n_steps = n_steps; p.random_policy = 0; p.act_stride = h->n; p.done_count = done_count; p.actions_out = actions_out;
    p.policy_w = weights; p.policy_b = bias; p.policy_actions = n_policy_actions;
    return launch_step(h, p, (cudaStream_t)stream);
}

int ngw_observe(ngw_handle* h, int32_t* obs, void* stream) {
    if (!h) return fail("null handle");
    if (h->obs_dim == 0) return 0;
    if (!obs || ((uintptr_t)obs & 15)) return fail("ngw_observe: obs must be a 16-byte aligned device pointer");
    CK(cudaSetDevice(h->device));
    return launch_step(h, step_params(h, nullptr, obs, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, h->n),
                       (cudaStream_t)stream);
}

static int ensure_host_path(ngw_handle* h) {
    if (h->hs[0]) return 0;
    for (int i = 0; i < HOST_STREAMS; i++) CK(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
    CK(cudaMalloc(&h->h_actions, (size_t)h->np * 4));
    if (h->obs_dim > 0) CK(cudaMalloc(&h->h_obs, (size_t)h->np * h->obs_dim * 4));
    // reward | step_cost | done | result share one allocation, n-element sections, so that a caller whose host buffers
    // have the same layout gets them with ONE device-to-host copy
    unsigned char* small = nullptr;
    CK(cudaMalloc(&small, (size_t)h->n * 10 + 64));
    h->h_reward = reinterpret_cast<float*>(small);
    h->h_cost = reinterpret_cast<float*>(small + (size_t)h->n * 4);
    h->h_done = small + (size_t)h->n * 8;
    h->h_result = small + (size_t)h->n * 9;
    return 0;
}

int ngw_set_message_buffer(ngw_handle* h, uint16_t* msg_dev) {
    if (!h) return fail("null handle");
    h->msg = msg_dev;
    return 0;
}

int ngw_step_host_end(ngw_handle* h) {
    if (!h) return fail("null handle");
    if (!h->hs[0]) return 0;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->hs[0]));
    return 0;
}

int ngw_step_host(ngw_handle* h, const int32_t* actions, int32_t* obs, float* reward, uint8_t* done, float* step_cost,
                  uint8_t* result, int32_t auto_reset, int32_t max_episode_steps) {
    if (ngw_step_host_begin(h, actions, obs, reward, done, step_cost, result, auto_reset, max_episode_steps)) return 1;
    return ngw_step_host_end(h);
}

int ngw_step_host_begin(ngw_handle* h, const int32_t* actions, int32_t* obs, float* reward, uint8_t* done,
                        float* step_cost, uint8_t* result, int32_t auto_reset, int32_t max_episode_steps) {
    if (!h) return fail("null handle");
    if (!actions || !reward || !done || !step_cost || !result) return fail("ngw_step_host: null pointer");
    CK(cudaSetDevice(h->device));
    if (ensure_host_path(h)) return 1;
    // The step kernel is ~1% of the PCIe time of its own outputs (17 MB of observations per 65,536 envs at ~55 GB/s),
    // so chunked compute/copy overlap buys nothing: one stream, one H2D, one launch, five D2H, one synchronize.
    long long n = h->n;
    cudaStream_t s = h->hs[0];
    size_t cnt = (size_t)n;
    CK(cudaMemcpyAsync(h->h_actions, actions, cnt * 4, cudaMemcpyHostToDevice, s));
    if (launch_step(h, step_params(h, h->h_actions, h->h_obs, h->h_reward, h->h_done, h->h_cost, h->h_result,
                                   auto_reset, max_episode_steps, 0, n), s)) return 1;
    if (h->obs_dim > 0 && obs)
        CK(cudaMemcpyAsync(obs, h->h_obs, cnt * h->obs_dim * 4, cudaMemcpyDeviceToHost, s));
    const unsigned char* r8 = reinterpret_cast<const unsigned char*>(reward);
    if (reinterpret_cast<const unsigned char*>(step_cost) == r8 + cnt * 4 && done == r8 + cnt * 8 && result == r8 + cnt * 9) {
        CK(cudaMemcpyAsync(reward, h->h_reward, cnt * 10, cudaMemcpyDeviceToHost, s));   // same layout: one copy
    } else {
        CK(cudaMemcpyAsync(reward, h->h_reward, cnt * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(done, h->h_done, cnt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(step_cost, h->h_cost, cnt * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(result, h->h_result, cnt, cudaMemcpyDeviceToHost, s));
    }
    return 0;
}

int ngw_agent_map(ngw_handle* h, int8_t* out, int32_t view, void* stream) {
    if (!h || !out || view < 1 || view > 32) return fail("ngw_agent_map: bad arguments");
    CK(cudaSetDevice(h->device));
    long long total = h->n * (2 * view + 1) * (2 * view + 1);
    agent_map_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->map, h->pose, out, h->n, h->ms,
                                                                                         view);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int ngw_stats(ngw_handle* h, double* out8_dev, int32_t reset_after, void* stream) {
    if (!h || !out8_dev) return fail("ngw_stats: null");
    CK(cudaSetDevice(h->device));
    stats_fold_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->stats, out8_dev, reset_after);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

int64_t ngw_launch_count(ngw_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
