// ngw_kernels.cu — kernels + C-ABI of libngw_b200.so (sm_100a only; see include/ngw.h for the contract).
//
// Data layout in HBM (struct of arrays, rows padded to a multiple of 32 envs so a warp's tile of 32 envs is
// one contiguous, 128-byte aligned span in every array):
//     map   int8  [Np][ms*ms]      pose uchar4 [Np] (row, col, facing, selected)      inventory int32 [Np][Is]
//     cfg_id uint8 [Np]            episode u32 [Np]     ep_len i32 [Np]     error_flags u32 [Np]
//
// step_kernel: one warp = one tile of 32 envs, one lane = one env.  The tile's grid rows and inventory rows are
// brought into shared memory with two TMA 1-D bulk copies (cp.async.bulk + mbarrier), the lanes run the
// flattened reference step on their row, cast the LidarInFront beams into an observation tile in shared
// memory, and the inventory tile and observation tile leave with two TMA bulk stores; pose / reward / done /
// step_cost / result are plain coalesced accesses.  Algorithmic bytes per env-step are in DESIGN.md.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ngw_device.cuh"

namespace ngw {

// ------------------------------------------------------------------ PTX helpers (TMA 1-D bulk copies, mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 cache-hinted variants (createpolicy + .L2::cache_hint)
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait sleeps in hardware; the time bound turns a lost transaction into a trap instead of a hung GPU
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ parameters
struct StepParams {
    const DevConfig* dcfgs;     // global-memory copy of the configs (cold paths, and the hot path when NC == 0)
    int8_t* map;
    uchar4* pose;
    int32_t* inv;
    const uint8_t* cfg_id;
    uint32_t* episode;
    int32_t* ep_len;
    uint32_t* err;
    const int32_t* actions;     // nullptr => observe only (no step, no outputs but obs)
    int32_t* obs;               // nullptr => no observation
    float* reward;
    uint8_t* done;
    float* cost;
    uint8_t* result;
    double* stats;              // [NGW_STAT_SLOTS][NGW_STAT_COUNT] or nullptr
    long long env_begin, env_end;   // env range of this launch (env_begin multiple of 32)
    long long first_gid;
    unsigned long long seed;
    int ms, cells, inv_stride, obs_dim;
    int map_bytes, inv_bytes, obs_bytes, region_bytes;   // per-warp shared-memory carve-up
    int auto_reset, max_episode_steps;
    int lidar_uniform;          // every config has the same beam tables (then config 0's are read, warp-uniformly)
    int cache_hints;            // bit 0: state tiles are loaded L2::evict_first, bit 1: the observation tile is stored evict_first
    int plain_store;            // 1 => write tiles back with ordinary coalesced stores instead of TMA bulk stores
    // K-step rollout (n_steps > 1 or random policy): the tile stays in shared memory across the steps
    int n_steps;                // steps per launch (1 for ngw_step)
    int random_policy;          // 1 => actions drawn on the device (Philox), `actions` is only a non-null marker
    long long act_stride;       // elements between consecutive steps in `actions` / `actions_out`
    unsigned long long policy_seed;
    int32_t* done_count;        // optional: episodes finished per env during the launch
    int32_t* actions_out;       // optional: actions taken
    const int32_t* policy_w;    // closed-loop linear policy: int32 [obs_dim][policy_actions] (nullptr = off)
    const int32_t* policy_b;    // int32 [policy_actions]
    int policy_actions;
    uint16_t* msg;              // optional: info['message'] codes (enum ngw_msg | arg << 5)
    int32_t* reset_list;        // single-step auto-reset: finished envs are queued here ...
    int32_t* reset_count;       // ... and regenerated by reset_list_kernel right after this launch
};

// Kernel argument block: the parameters plus up to NC configs INLINE, so that every config read in the hot path
// is a constant-bank operand (c[0x0][..]) instead of a global load.  NC == 0 falls back to global memory.
template <int NC>
struct StepArgs {
    StepParams p;
    DevConfig cfg[NC > 0 ? NC : 1];
};

#define NGW_STAT_SLOTS 512

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

__device__ __forceinline__ void warp_copy16(void* dst, const void* src, int bytes, int lane) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = lane; i < (bytes >> 4); i += 32) d[i] = s[i];
}

// Cold: fused auto-reset.  Called by the whole step warp; every lane that finished an episode is regenerated in turn by
// all 32 lanes (reset_env_warp), then its grid row goes back to HBM with coalesced stores.
__device__ __noinline__ void auto_reset_warp(const StepParams& p, const DevConfig* dcfgs, int cfg_i, bool need,
                                             int8_t* smap, int32_t* sinv, uint32_t* hist, long long e0, uchar4& ps) {
    const int lane = threadIdx.x & 31;
    __syncwarp();                                                    // every lane's step writes to the tile are visible
    uint32_t pending = __ballot_sync(0xFFFFFFFFu, need);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const long long e = e0 + src;
        const int ci = __shfl_sync(0xFFFFFFFFu, cfg_i, src);
        uint32_t ep = p.episode[e] + 1;
        __syncwarp();
        if (lane == 0) p.episode[e] = ep;
        int8_t* m = smap + src * p.cells;
        int32_t* inv = sinv + src * p.inv_stride;
        int r = 0, c = 0, f = 0, sel = 0;
        uint32_t err = reset_env_warp(&dcfgs[ci].c, m, inv, p.ms, p.inv_stride, p.seed, (uint64_t)(p.first_gid + e), ep,
                                      true, 0, NGW_MAX_RESET_OPS, hist, r, c, f, sel);
        int8_t* grow = p.map + e * p.cells;
        for (int i = lane; i < p.cells; i += 32) grow[i] = m[i];
        if (lane == src) {
            ps = make_uchar4((unsigned char)r, (unsigned char)c, (unsigned char)f, (unsigned char)sel);
            if (err) p.err[e] |= err;
        }
        __syncwarp();
    }
}

// Episode statistics of one tile: warp reductions + one atomic per counter into one of NGW_STAT_SLOTS slots.
__device__ __forceinline__ void tile_stats(double* stats, int lane, int valid, int done, int success, int did_reset,
                                           int invalid, int reward, float cost) {
    // five small counts (each <= 32) share one reduction: 6 bits apiece
    unsigned packed = (unsigned)(valid ? done : 0) | ((unsigned)success << 6) | ((unsigned)did_reset << 12) |
                      ((unsigned)invalid << 18) | ((unsigned)(valid ? 1 : 0) << 24);
    packed = __reduce_add_sync(0xFFFFFFFFu, packed);
    const int r_sum = __reduce_add_sync(0xFFFFFFFFu, reward);
    const float c_sum = warp_sum(cost);
    if (lane == 0) {
        const int n_done = packed & 63, n_succ = (packed >> 6) & 63, n_reset = (packed >> 12) & 63;
        const int n_inv = (packed >> 18) & 63, n_valid = (packed >> 24) & 63;
        double* s = stats + (size_t)(blockIdx.x % NGW_STAT_SLOTS) * NGW_STAT_COUNT;
        atomicAdd(&s[NGW_STAT_STEPS], (double)(n_valid - n_inv));
        atomicAdd(&s[NGW_STAT_REWARD_SUM], (double)r_sum);
        atomicAdd(&s[NGW_STAT_COST_SUM], (double)c_sum);
        if (n_done) atomicAdd(&s[NGW_STAT_EPISODES], (double)n_done);
        if (n_succ) atomicAdd(&s[NGW_STAT_SUCCESSES], (double)n_succ);
        if (n_reset) atomicAdd(&s[NGW_STAT_RESETS], (double)n_reset);
        if (n_inv) atomicAdd(&s[NGW_STAT_INVALID], (double)n_inv);
    }
}

// ------------------------------------------------------------------ the fused step + LidarInFront kernel
// One CTA = one tile of 32 consecutive envs, G = blockDim.x / 32 warps.  Lane l of every warp owns env l of the tile.
// Warp 0 runs the flattened step; then all G warps cast 8/G lidar beams each for their lane's env.  G = 1 is the plain
// one-warp-per-tile kernel; G > 1 shortens the per-tile latency where shared memory limits the tiles per SM.
template <bool kTma, int NC, bool kMulti>
__global__ void __launch_bounds__(256) step_kernel(const __grid_constant__ StepArgs<NC> args) {
    extern __shared__ __align__(128) unsigned char smem[];
    const StepParams& p = args.p;
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5, G = blockDim.x >> 5;
    const long long e0 = p.env_begin + (long long)blockIdx.x * 32;
    const long long e = e0 + lane;
    const bool valid = e < p.env_end;
    const bool full_tile = e0 + 32 <= p.env_end;
    const bool stepping = p.actions != nullptr;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    int8_t* szero = reinterpret_cast<int8_t*>(smem + 8);             // 8 bytes that always read 0 (landed lidar beams park here)
    uchar4* spose = reinterpret_cast<uchar4*>(smem + 16);            // pose after the step, for the other warps
    constexpr int kScratch = kMulti ? 1024 : 0;                      // radix-select histogram of the in-place auto-reset (rollout only)
    uint32_t* sscratch = reinterpret_cast<uint32_t*>(smem + 16 + 128);
    int8_t* smap = reinterpret_cast<int8_t*>(smem + 16 + 128 + kScratch);
    int32_t* sinv = reinterpret_cast<int32_t*>(smem + 16 + 128 + kScratch + p.map_bytes);
    int32_t* sobs = reinterpret_cast<int32_t*>(smem + 16 + 128 + kScratch + p.map_bytes + p.inv_bytes);

    // ---- prologue without global memory: barrier, zero pad, zeroed observation tile
    const int8_t* gmap = p.map + e0 * p.cells;
    int32_t* ginv = p.inv + e0 * p.inv_stride;
    if (threadIdx.x == 0) {
        if (kTma) mbar_init(bar, 1);
        *reinterpret_cast<uint64_t*>(szero) = 0ull;
    }
    if (p.obs != nullptr) {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* o4 = reinterpret_cast<uint4*>(sobs);
        for (int i = threadIdx.x; i < (p.obs_bytes >> 4); i += blockDim.x) o4[i] = z;
    }
    __syncthreads();                                                 // barrier init visible before anyone waits on it
    asm volatile("griddepcontrol.wait;" ::: "memory");               // previous kernel of the stream done + visible

    // ---- stage the tile: grid rows + inventory rows (state arrays are padded, a full tile is always readable)
    if (kTma) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (uint32_t)(p.map_bytes + p.inv_bytes));
            if (p.cache_hints & 1) {
                uint64_t pol = policy_evict_first();
                bulk_g2s_hint(smap, gmap, (uint32_t)p.map_bytes, bar, pol);
                bulk_g2s_hint(sinv, ginv, (uint32_t)p.inv_bytes, bar, pol);
            } else {
                bulk_g2s(smap, gmap, (uint32_t)p.map_bytes, bar);
                bulk_g2s(sinv, ginv, (uint32_t)p.inv_bytes, bar);
            }
        }
    } else {
        const uint4* s4 = reinterpret_cast<const uint4*>(gmap);
        uint4* d4 = reinterpret_cast<uint4*>(smap);
        for (int i = threadIdx.x; i < (p.map_bytes >> 4); i += blockDim.x) d4[i] = s4[i];
        s4 = reinterpret_cast<const uint4*>(ginv);
        d4 = reinterpret_cast<uint4*>(sinv);
        for (int i = threadIdx.x; i < (p.inv_bytes >> 4); i += blockDim.x) d4[i] = s4[i];
    }

    // ---- while the copies fly: per-lane scalars
    const int cfg_i = (NC == 1) ? 0 : (int)p.cfg_id[e];
    const DevConfig& dc = (NC == 1) ? args.cfg[0] : (NC > 1 ? args.cfg[cfg_i] : p.dcfgs[cfg_i]);
    const ngw_config& cfg = dc.c;
    uchar4 ps = make_uchar4(0, 0, 0, 0);
    int action = 0;
    if (g == 0) {
        ps = p.pose[e];
        const bool given_actions = !(kMulti && (p.random_policy || p.policy_w != nullptr));   // else `actions` is a marker
        if (stepping && valid && given_actions) action = p.actions[e];
    }

    if (kTma) mbar_wait(bar, 0);
    else __syncthreads();

    StepOut st_out;                                                  // one-step kernel: statistics are folded after the lidar,
    st_out.reward = 0; st_out.done = 0; st_out.result = 0; st_out.cost = 0.0f; st_out.msg = 0;   // off the path to the barrier
    int st_success = 0, st_reset = 0, st_invalid = 0;
    EnvRow env;
    env.m = smap + lane * p.cells;
    env.gm = p.map + e * p.cells;
    env.inv = sinv + lane * p.inv_stride;
    env.ms = p.ms;

    if (g == 0) {
        env.r = ps.x; env.c = ps.y; env.facing = ps.z; env.sel = ps.w;
        if (stepping) {
            StepOut o;
            o.reward = 0; o.done = 0; o.result = 0; o.cost = 0.0f; o.msg = 0;
            float reward_sum = 0.0f, cost_sum = 0.0f;
            int done_count = 0;
            const int n_steps = kMulti ? p.n_steps : 1;               // kMulti == false: the plain one-step kernel
            const bool random_policy = kMulti && p.random_policy;
            for (int t = 0; t < n_steps; t++) {
                int next_action = 0;                                  // prefetch the next step's action behind this step
                if (!random_policy && !(kMulti && p.policy_w != nullptr) && t + 1 < n_steps && valid)
                    next_action = p.actions[(t + 1) * p.act_stride + e];
                if (kMulti && p.policy_w != nullptr) {                // closed loop: observe, then greedy linear policy
                    int32_t* row = sobs + lane * p.obs_dim;
                    for (int j = 0; j < p.obs_dim; j++) row[j] = 0;
                    if (valid && cfg.n_beams > 0) lidar_observe(env, dc, dc.lidar, row, szero, 0, 1, true);
                    if (valid) {
                        int acc[16];
                        const int A = p.policy_actions;
#pragma unroll
                        for (int a = 0; a < 16; a++) acc[a] = a < A ? p.policy_b[a] : 0;
                        const int D = cfg.n_lidar_items * cfg.n_beams + cfg.n_inv_obs;
                        for (int j = 0; j < D; j++) {
                            const int v = row[j];
                            if (v == 0) continue;                      // the observation is sparse (<= 8 hits + inventory)
                            const int32_t* w = p.policy_w + (size_t)j * A;
#pragma unroll
                            for (int a = 0; a < 16; a++) if (a < A) acc[a] += v * w[a];
                        }
                        int best = 0, best_v = acc[0];
                        const int n_valid_actions = cfg.n_actions < A ? cfg.n_actions : A;
#pragma unroll
                        for (int a = 1; a < 16; a++) if (a < n_valid_actions && acc[a] > best_v) { best_v = acc[a]; best = a; }
                        action = best;
                    }
                } else if (random_policy && valid) {                  // uniform over the config's action ids
                    Philox pr;
                    pr.init(p.policy_seed, (uint64_t)(p.first_gid + e), (uint32_t)t, 0xFFF);
                    action = (int)pr.below((uint32_t)(cfg.n_actions > 0 ? cfg.n_actions : 1));
                }
                if (kMulti && p.actions_out != nullptr && valid) p.actions_out[t * p.act_stride + e] = action;
                o.reward = 0; o.done = 0; o.result = 0; o.cost = 0.0f; o.msg = 0;
                int invalid = 0, did_reset = 0, success = 0;
                if (valid) {
                    ngw_action_entry a;
                    a.op = NGW_OP_INVALID;
                    if (action >= 0 && action < cfg.n_actions) {
                        uint2 raw = *reinterpret_cast<const uint2*>(&cfg.actions[action]);
                        memcpy(&a, &raw, sizeof(a));
                    }
                    if (a.op == NGW_OP_INVALID) {                     // wrappers.py:76 / pogostick_v1_env.py:236 would raise
                        invalid = 1;
                        p.err[e] |= NGW_ERR_INVALID_ACTION;
                    } else {
                        step_env(env, cfg, a, o);
                        success = o.done && env.inv[cfg.id_goal] >= 1;
                        int finished = o.done;
                        if (p.max_episode_steps > 0) {
                            int len = p.ep_len[e] + 1;
                            if (len >= p.max_episode_steps) { finished = 1; o.done = 1; } // harness truncation knob
                            p.ep_len[e] = finished && p.auto_reset ? 0 : len;
                        }
                        if (finished && p.auto_reset) { did_reset = 1; }
                    }
                    ps = make_uchar4((unsigned char)env.r, (unsigned char)env.c, (unsigned char)env.facing,
                                     (unsigned char)env.sel);
                    reward_sum += (float)o.reward; cost_sum += o.cost; done_count += o.done;
                }
                if (p.auto_reset) {
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, did_reset);
                    if (bal != 0 && !kMulti) {
                        // single step: queue the finished envs; reset_list_kernel (next in the stream, one warp per env at
                        // full occupancy) regenerates them and overwrites their observation rows
                        int base = 0;
                        if (lane == 0) base = atomicAdd(p.reset_count, __popc(bal));
                        base = __shfl_sync(0xFFFFFFFFu, base, 0);
                        if (did_reset) p.reset_list[base + __popc(bal & ((1u << lane) - 1u))] = (int)e;
                    } else if (bal != 0) {
                        // rollout: the next step needs the new episode now -> regenerate in place, warp-cooperatively
                        auto_reset_warp(p, p.dcfgs, cfg_i, did_reset != 0, smap, sinv, sscratch, e0, ps);
                        env.r = ps.x; env.c = ps.y; env.facing = ps.z; env.sel = ps.w;
                    }
                }
                if (kMulti && p.stats != nullptr)
                    tile_stats(p.stats, lane, valid, o.done, success, did_reset, invalid, o.reward, o.cost);
                if (!kMulti) { st_success = success; st_reset = did_reset; st_invalid = invalid; }
                action = next_action;
            }
            if (!kMulti) st_out = o;                                  // one-step kernel: outputs are stored after the lidar
            if (kMulti && valid) {
                p.pose[e] = ps;
                p.reward[e] = reward_sum;
                p.done[e] = (uint8_t)o.done;
                p.cost[e] = cost_sum;
                p.result[e] = (uint8_t)o.result;
                if (p.done_count != nullptr) p.done_count[e] = done_count;
                if (p.msg != nullptr) p.msg[e] = (uint16_t)o.msg;
            }
        }
        if (G > 1) spose[lane] = ps;
    }
    if (G > 1) {
        __syncthreads();                                             // step results (grid, inventory, pose) visible to all warps
        ps = spose[lane];
        env.r = ps.x; env.c = ps.y; env.facing = ps.z; env.sel = ps.w;
    }

    // ---- LidarInFront observation of the (possibly auto-reset) state into the shared-memory tile
    if (p.obs != nullptr && valid && cfg.n_beams > 0)
        lidar_observe(env, dc, (NC > 1 && p.lidar_uniform) ? args.cfg[0].lidar : dc.lidar, sobs + lane * p.obs_dim, szero, g,
                      G, g == G - 1);

    if (!kMulti && g == 0 && stepping) {                             // outputs and statistics, off the path to the barrier
        if (valid) {
            p.pose[e] = ps;
            p.reward[e] = (float)st_out.reward;
            p.done[e] = (uint8_t)st_out.done;
            p.cost[e] = st_out.cost;
            p.result[e] = (uint8_t)st_out.result;
            if (p.msg != nullptr) p.msg[e] = (uint16_t)st_out.msg;
        }
        if (p.stats != nullptr)
            tile_stats(p.stats, lane, valid, st_out.done, st_success, st_reset, st_invalid, st_out.reward, st_out.cost);
    }

    // Programmatic dependent launch: this tile's compute is done, let the next kernel of the stream start scheduling its
    // CTAs; its prologue (up to griddepcontrol.wait) touches no global memory, so it overlaps this kernel's store phase.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- write back: inventory tile (only when stepping) and observation tile
    __syncthreads();
    if (kTma && full_tile && !p.plain_store) {
        if (threadIdx.x < 32) {
            fence_async_smem();
            __syncwarp();
            if (threadIdx.x == 0) {
                if (p.cache_hints & 2) {
                    uint64_t pol = policy_evict_first();
                    if (stepping) bulk_s2g(ginv, sinv, (uint32_t)p.inv_bytes);
                    if (p.obs != nullptr) bulk_s2g_hint(p.obs + e0 * p.obs_dim, sobs, (uint32_t)p.obs_bytes, pol);
                } else {
                    if (stepping) bulk_s2g(ginv, sinv, (uint32_t)p.inv_bytes);
                    if (p.obs != nullptr) bulk_s2g(p.obs + e0 * p.obs_dim, sobs, (uint32_t)p.obs_bytes);
                }
                bulk_commit();
                bulk_wait_read0();                                   // shared memory must outlive the reads
            }
        }
    } else {
        if (stepping) {
            const uint4* s4 = reinterpret_cast<const uint4*>(sinv);
            uint4* d4 = reinterpret_cast<uint4*>(ginv);
            for (int i = threadIdx.x; i < (p.inv_bytes >> 4); i += blockDim.x) d4[i] = s4[i];
        }
        if (p.obs != nullptr) {
            int n = (int)((p.env_end - e0 < 32 ? p.env_end - e0 : 32)) * p.obs_dim;
            int32_t* gobs = p.obs + e0 * p.obs_dim;
            for (int i = threadIdx.x; i < n; i += blockDim.x) gobs[i] = sobs[i];
        }
    }
}

// ------------------------------------------------------------------ cold-path kernels (one thread per env, global memory)
struct ResetParams {
    const DevConfig* dcfgs;
    int8_t* map;
    uchar4* pose;
    int32_t* inv;
    const uint8_t* cfg_id;
    uint32_t* episode;
    int32_t* ep_len;
    uint32_t* err;
    const uint8_t* mask;
    long long n_envs, first_gid;
    unsigned long long seed;
    int ms, cells, inv_stride;
    int phase;   // 0: base + ops before the reset observation, 1: ops after it, 2: everything
    const uint8_t* zero_byte;   // a global byte that always reads 0 (see lidar_observe)
    const int32_t* reset_list;  // reset_list_kernel: queue written by the step kernel
    int32_t* reset_count;
    int32_t* done_ctas;
    int32_t* obs;
    int obs_dim;
};

#define NGW_RESET_WARPS 4
// The cold kernels regenerate an env on a shared-memory copy of its rows (every pass of reset_env_warp would otherwise
// pay HBM latency) and write the rows back with coalesced stores.
struct ResetScratch {
    uint32_t hist[256];
    int32_t inv[NGW_MAX_ITEMS];
    int8_t row[NGW_MAX_MAP_SIZE * NGW_MAX_MAP_SIZE];
};

__device__ __forceinline__ void rows_to_smem(ResetScratch& sc, const int8_t* m, const int32_t* inv, int cells, int inv_stride,
                                             int lane) {
    for (int i = lane; i < cells; i += 32) sc.row[i] = m[i];
    for (int i = lane; i < inv_stride; i += 32) sc.inv[i] = inv[i];
    __syncwarp();
}
__device__ __forceinline__ void rows_from_smem(const ResetScratch& sc, int8_t* m, int32_t* inv, int cells, int inv_stride,
                                               int lane) {
    __syncwarp();
    for (int i = lane; i < cells; i += 32) m[i] = sc.row[i];
    for (int i = lane; i < inv_stride; i += 32) inv[i] = sc.inv[i];
}

__global__ void __launch_bounds__(32 * NGW_RESET_WARPS) reset_kernel(const ResetParams p) {
    __shared__ ResetScratch scratch[NGW_RESET_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long e = (long long)blockIdx.x * NGW_RESET_WARPS + warp;       // one warp regenerates one env
    if (e >= p.n_envs) return;
    if (p.mask != nullptr && p.mask[e] == 0) return;
    ResetScratch& sc = scratch[warp];
    const ngw_config* cfg = &p.dcfgs[p.cfg_id[e]].c;
    uchar4 ps = p.pose[e];
    int r = ps.x, c = ps.y, f = ps.z, sel = ps.w;
    uint64_t gid = (uint64_t)(p.first_gid + e);
    int k = cfg->reset_obs_after_ops;
    int8_t* m = p.map + e * p.cells;
    int32_t* inv = p.inv + e * p.inv_stride;
    if (p.phase != 1) {
        uint32_t ep = p.episode[e] + 1;
        __syncwarp();
        uint32_t err = reset_env_warp(cfg, sc.row, sc.inv, p.ms, p.inv_stride, p.seed, gid, ep, true, 0,
                                      p.phase == 0 ? k : NGW_MAX_RESET_OPS, sc.hist, r, c, f, sel);
        if (lane == 0) { p.episode[e] = ep; p.ep_len[e] = 0; p.err[e] = err; }
    } else {
        uint32_t ep = p.episode[e];
        rows_to_smem(sc, m, inv, p.cells, p.inv_stride, lane);
        reset_env_warp(cfg, sc.row, sc.inv, p.ms, p.inv_stride, p.seed, gid, ep, false, k, NGW_MAX_RESET_OPS, sc.hist, r, c,
                       f, sel);
    }
    rows_from_smem(sc, m, inv, p.cells, p.inv_stride, lane);
    if (lane == 0) p.pose[e] = make_uchar4((unsigned char)r, (unsigned char)c, (unsigned char)f, (unsigned char)sel);
}

// Second half of the single-step auto-reset: a grid-stride loop of warps over the queue the step kernel filled.  Each
// warp regenerates one env in HBM (reset_env_warp), then 8 lanes cast its LidarInFront beams into its observation row.
// The last CTA to finish empties the queue for the next step.
__global__ void __launch_bounds__(32 * NGW_RESET_WARPS) reset_list_kernel(const ResetParams p) {
    __shared__ ResetScratch scratch[NGW_RESET_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ResetScratch& sc = scratch[warp];
    const int count = *reinterpret_cast<volatile int32_t*>(p.reset_count);
    for (int i = blockIdx.x * NGW_RESET_WARPS + warp; i < count; i += gridDim.x * NGW_RESET_WARPS) {
        const long long e = p.reset_list[i];
        const DevConfig& dc = p.dcfgs[p.cfg_id[e]];
        uchar4 ps = p.pose[e];
        int r = ps.x, c = ps.y, f = ps.z, sel = ps.w;
        int8_t* m = p.map + e * p.cells;
        int32_t* inv = p.inv + e * p.inv_stride;
        uint32_t ep = p.episode[e] + 1;
        __syncwarp();
        uint32_t err = reset_env_warp(&dc.c, sc.row, sc.inv, p.ms, p.inv_stride, p.seed, (uint64_t)(p.first_gid + e), ep,
                                      true, 0, NGW_MAX_RESET_OPS, sc.hist, r, c, f, sel);
        rows_from_smem(sc, m, inv, p.cells, p.inv_stride, lane);
        if (lane == 0) {
            p.episode[e] = ep;
            p.pose[e] = make_uchar4((unsigned char)r, (unsigned char)c, (unsigned char)f, (unsigned char)sel);
            if (err) p.err[e] |= err;
        }
        if (p.obs != nullptr) {                                       // observation of the new episode replaces the row
            int32_t* row = p.obs + e * p.obs_dim;
            for (int k = lane; k < p.obs_dim; k += 32) row[k] = 0;
            __syncwarp();
            EnvRow env;                                                   // lidar on the shared-memory copy
            env.m = sc.row; env.gm = nullptr; env.inv = sc.inv; env.ms = p.ms;
            env.r = r; env.c = c; env.facing = f; env.sel = sel;
            if (lane == 0) sc.hist[0] = 0;                                // a shared-memory byte that reads 0
            __syncwarp();
            const int8_t* zero = reinterpret_cast<const int8_t*>(sc.hist);
            if (dc.c.n_beams > 0) {
                if (dc.lidar.fast) { if (lane < 8) lidar_observe(env, dc, dc.lidar, row, zero, lane, 8, lane == 7); }
                else if (lane == 0) lidar_observe(env, dc, dc.lidar, row, zero, 0, 1, true);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.done_ctas, 1) == (int)gridDim.x - 1) { *p.reset_count = 0; *p.done_ctas = 0; }
    }
}

__global__ void observe_masked_kernel(const ResetParams p, int32_t* obs, int obs_dim) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_envs) return;
    if (p.mask != nullptr && p.mask[e] == 0) return;
    const DevConfig& dc = p.dcfgs[p.cfg_id[e]];
    EnvRow env;
    env.m = p.map + e * p.cells;
    env.gm = nullptr;
    env.inv = p.inv + e * p.inv_stride;
    env.ms = p.ms;
    uchar4 ps = p.pose[e];
    env.r = ps.x; env.c = ps.y; env.facing = ps.z; env.sel = ps.w;
    int32_t* row = obs + e * obs_dim;
    for (int i = 0; i < obs_dim; i++) row[i] = 0;
    if (dc.c.n_beams > 0) lidar_observe(env, dc, dc.lidar, row, reinterpret_cast<const int8_t*>(p.zero_byte), 0, 1, true);
}

// AgentMap.get_agentView (observation_wrappers.py:98-118): zero-padded (2v+1)^2 crop centred on the agent
__global__ void agent_map_kernel(const int8_t* map, const uchar4* pose, int8_t* out, long long n, int ms, int view) {
    const int side = 2 * view + 1;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * side * side) return;
    long long e = idx / (side * side);
    int k = (int)(idx - e * side * side);
    int r = pose[e].x - view + k / side, c = pose[e].y - view + k % side;
    out[idx] = (r >= 0 && r < ms && c >= 0 && c < ms) ? map[e * ms * ms + r * ms + c] : (int8_t)0;
}

__global__ void set_cfg_kernel(const int32_t* src, uint8_t* dst, long long n, int n_cfgs, uint32_t* err) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int v = src ? src[e] : 0;
    if (v < 0 || v >= n_cfgs) { v = 0; err[e] |= 0x80000000u; }
    dst[e] = (uint8_t)v;
}

__global__ void stats_fold_kernel(double* slots, double* out, int reset_after) {
    int k = threadIdx.x;
    if (k >= NGW_STAT_COUNT) return;
    double s = 0.0;
    for (int i = 0; i < NGW_STAT_SLOTS; i++) {
        s += slots[i * NGW_STAT_COUNT + k];
        if (reset_after) slots[i * NGW_STAT_COUNT + k] = 0.0;
    }
    out[k] = s;
}

}  // namespace ngw

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
