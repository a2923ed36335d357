#!/usr/bin/env python
"""bench.py — env-steps/sec of the fused step + LidarInFront hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, weak scaling)

A "step" is ONE pass of the hot path over ONE batch of 65,536 envs of BASELINE config C2
(NovelGridworld-Pogostick-v1 + LimitActions(10 actions) + LidarInFront(8 beams)), random actions.
To keep L2 cold between timed iterations the run rotates over several independent batches whose combined
working set (state + outputs) exceeds the 126 MB L2 ("inputs larger than L2").

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM (CUDA-graph
replay of the K launches, no host in the loop); `e2e` = the same metric through the host-buffer C-ABI call
(ngw_step_host: pinned H2D of actions, kernel, D2H of obs/reward/done/step_cost/result every step);
`roofline` = algorithmic bytes per launch / average launch duration vs the measured HBM copy peak;
`cpu_baseline` = the C oracle port on the box's host cores.  `--impl reference` times that CPU port alone.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
    os.environ['NCCL_DEBUG'] = 'WARN'                  # keep NCCL's version banner off stdout: ONE JSON line only

ENVS_PER_BATCH = 65536
C2_SET = ['Forward', 'Left', 'Right', 'Break', 'Place_tree_tap', 'Extract_rubber',
          'Craft_plank', 'Craft_stick', 'Craft_tree_tap', 'Craft_pogo_stick']
METRIC = "env-steps/sec, Pogostick-v1+LidarInFront, 1/2/4/8 B200 vs host-CPU reference"
N_ACTION_SETS = 16
GRAPH_STEPS = 1024


def build_c2_chain(num_envs=1, device=None):
    import gym_novel_gridworlds_b200 as gym
    env = gym.make('NovelGridworld-Pogostick-v1', num_envs=num_envs, device=device)
    env = gym.LimitActions(env, set(C2_SET))
    return gym.LidarInFront(env, num_beams=8)


def build_workload(name):
    """BASELINE.json configs -> (description, [compiled configs], envs per batch per GPU, cfg-id rule, step kwargs).
    C2 is the headline (default); C3-C5 are extra measurements (`--workload`)."""
    import numpy as np
    import gym_novel_gridworlds_b200 as gym
    from gym_novel_gridworlds_b200.compiler import compile_chain

    def pogo(extra_actions=(), map_size=10):
        env = gym.make('NovelGridworld-Pogostick-v1')
        env.map_size = map_size
        env = gym.LimitActions(env, set(C2_SET) | set(extra_actions))
        return gym.LidarInFront(env, num_beams=8)

    if name == 'C2':
        return ("C2: NovelGridworld-Pogostick-v1 + LimitActions(10) + LidarInFront(8), 65536 envs/batch, "
                "uniform random actions", [compile_chain(pogo())], 65536, None, {})
    if name == 'C3':
        env = gym.make('NovelGridworld-Bow-v1')
        env = gym.LidarInFront(env, num_beams=8)
        env = gym.inject_novelty(env, 'axe', 'medium', 'wooden', '')
        env = gym.inject_novelty(env, 'fence', 'hard', 'oak', '')
        return ("C3: NovelGridworld-Bow-v1 + LidarInFront(8) + axe(medium, wooden) + fence(hard, oak), 262144 envs/batch",
                [compile_chain(env)], 262144, None, {})
    if name in ('C4', 'C4-blocked'):
        np.random.seed(4)
        chains = [gym.inject_novelty(pogo(['Chop']), 'addchop'), gym.inject_novelty(pogo(['Jump']), 'addjump'),
                  gym.inject_novelty(pogo(), 'additem', 'medium', 'spring'),
                  gym.inject_novelty(pogo(), 'remapaction', 'hard')]
        rule = 'interleaved' if name == 'C4' else 'blocked'
        return ("C4: Pogostick-v1 + LimitActions + LidarInFront(8), per-env novelty addchop / addjump / additem(medium) / "
                "remapaction(hard), config ids %s, 1048576 envs/batch, one launch per step" % rule,
                [compile_chain(c) for c in chains], 1048576, rule, {})
    if name == 'C5':
        env = gym.inject_novelty(pogo(map_size=40), 'additem', 'hard', 'spring')
        return ("C5: Pogostick-v1 map_size 40 + additem(hard) + LidarInFront(8), 524288 envs/batch per GPU, fused "
                "auto-reset, max_episode_steps 256 (harness truncation knob)",
                [compile_chain(env)], 524288, None, {'auto_reset': True, 'max_episode_steps': 256})
    raise ValueError(name)


def algorithmic_bytes_per_env_step(cc):
    """read ms^2 + 4 (pose) + 4 I (inventory) + 4 (action); write 4 (pose) + 4 I + 4 (L B + I_obs)
    + 4 (reward) + 4 (step_cost) + 1 (done) + 1 (result); grid write-back counted as 0."""
    ms, n_items, d = cc.map_size, cc.n_items, cc.obs_dim
    return (ms * ms + 4 + 4 * n_items + 4) + (4 + 4 * n_items + 4 * d + 4 + 4 + 1 + 1)


def measured_hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.max_mhz, self.ok = index, [], False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((time.perf_counter(), mhz, reasons, util))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self, t0, t1):
        names = {0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x10: 'sync_boost',
                 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
                 0x100: 'display_clock_setting'}
        if not self.samples:                            # NVML unavailable: one nvidia-smi reading as a last resort
            try:
                import subprocess
                q = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=clocks.sm,clocks.max.sm,'
                                    'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
                                    'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap',
                                    '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=20)
                f = [x.strip() for x in q.stdout.strip().split(',')]
                labels = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
                return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                        "reasons": [n for n, v in zip(labels, f[2:]) if v.lower().startswith('active')],
                        "samples": 1, "source": "nvidia-smi after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or [s for s in self.samples if s[3] > 0] or self.samples
        reasons = 0
        for s in inside:
            reasons |= s[2]
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in names.items() if reasons & b), "samples": len(inside)}


def cpu_port_run(cc, envs, steps, warmup, threads, time_budget_s=None):
    """Times the C oracle port (oracle/ngw_oracle.c) on `envs` environments with all host threads.  Threads own
    contiguous env slices and run them without a per-step barrier (envs are independent) — the fastest honest CPU
    arrangement of the reference algorithm; every step still writes obs/reward/done/step_cost/result."""
    from oracle.oracle_lib import OracleBatch
    ob = OracleBatch([cc], envs)
    ob.reset_legacy(1)
    rng = np.random.RandomState(1234)
    acts = np.stack([rng.randint(0, cc.c.n_actions, size=envs).astype(np.int32) for _ in range(N_ACTION_SETS)])
    if warmup:
        ob.rollout(acts, warmup, n_threads=threads)
    done, chunk = 0, 16
    t0 = time.perf_counter()
    while done < steps:
        n = min(chunk, steps - done)
        ob.rollout(acts, n, n_threads=threads)
        done += n
        if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
            break
    return done, time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself is Python and
    cannot travel to the GPU box, so this is its C port (the oracle) with every host thread.  A step is one pass
    over S envs of the same C2 workload, S bounded so that K steps end within a few minutes."""
    if rank != 0:
        return
    from gym_novel_gridworlds_b200.compiler import compile_chain
    cc = compile_chain(build_c2_chain())
    threads = os.cpu_count() or 1
    K = max(args.steps, 1)
    envs = ENVS_PER_BATCH
    while envs > 1024 and envs * K > 4e8:
        envs //= 2
    steps, dt = cpu_port_run(cc, envs, K, args.warmup, threads)
    value = steps * envs / dt
    sample = ("%d steps x %d envs of C2 (bounded sample of the 65536-env batch), C port of the reference path "
              "(oracle/ngw_oracle.c), %d threads, no per-step barrier" % (steps, envs, threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "C2: NovelGridworld-Pogostick-v1 + LimitActions(10) + LidarInFront(8), "
                               "uniform random actions", "envs_per_step": envs},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from gym_novel_gridworlds_b200.runtime import BatchHandle

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created (NCCL_DEBUG=VERSION is set in this
        # image); stdout must carry ONE JSON line, so fd 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    desc, compiled, envs, cfg_rule, step_kw = build_workload(args.workload)
    n_cfg = len(compiled)
    # algorithmic bytes: mean over the configs of the batch (+1 byte config id per env when mixed)
    bytes_step = float(np.mean([algorithmic_bytes_per_env_step(cc) for cc in compiled])) + (1 if n_cfg > 1 else 0)
    per_batch_ws = envs * bytes_step
    n_batches = max(2, int(np.ceil(1.6 * 126e6 / per_batch_ws)))       # combined working set >= 1.6x L2
    batches, act_sets = [], []
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    for b in range(n_batches):
        gid0 = (rank * n_batches + b) * envs
        cfg_id = None
        if n_cfg > 1:
            idx = np.arange(envs, dtype=np.int64)
            cfg_id = (idx % n_cfg) if cfg_rule == 'interleaved' else np.minimum(idx * n_cfg // envs, n_cfg - 1)
        h = BatchHandle(compiled, envs, device=dev, seed=0, first_env_gid=gid0, cfg_id=cfg_id)
        h.reset()
        if step_kw.get('max_episode_steps', 0) > 0:                     # stagger episode ages so truncation-resets spread evenly
            h.ep_len.copy_(torch.randint(0, step_kw['max_episode_steps'], (envs,), generator=gen, device=dev,
                                         dtype=torch.int32))
        batches.append(h)
    n_act = torch.tensor([cc.c.n_actions for cc in compiled], device=dev, dtype=torch.int64)
    per_env_n = n_act[batches[0].cfg_id.long()]
    n_sets = N_ACTION_SETS if envs <= 262144 else 4
    for _ in range(n_sets):
        r = torch.randint(0, 1 << 30, (envs,), generator=gen, device=dev, dtype=torch.int64)
        act_sets.append((r % per_env_n).to(torch.int32))
    flags = sum(int((h.error_flags != 0).sum().item()) for h in batches)

    def one_step(i):
        batches[i % n_batches].step(act_sets[i % n_sets], **step_kw)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- warm-up (eager launches), then capture the launch sequence into CUDA graphs
    W, K = max(args.warmup, 3), args.steps
    for i in range(W):
        one_step(i)
    torch.cuda.synchronize(dev)
    launches_before = sum(h.launch_count() for h in batches)
    g_steps = min(K, GRAPH_STEPS)
    g_steps -= g_steps % n_batches if g_steps >= n_batches else 0       # whole rotations per replay
    g_steps = max(g_steps, 1)
    stream = torch.cuda.Stream(dev)

    def capture(n, n_streams=1):
        g = torch.cuda.CUDAGraph()
        side = [torch.cuda.Stream(dev) for _ in range(n_streams)] if n_streams > 1 else []
        with torch.cuda.stream(stream):
            with torch.cuda.graph(g, stream=stream):
                if not side:
                    for i in range(n):
                        one_step(i)
                else:                                   # independent batches on parallel branches of the graph
                    for s_ in side:
                        s_.wait_stream(stream)
                    for i in range(n):
                        with torch.cuda.stream(side[(i % n_batches) % n_streams]):
                            one_step(i)
                    for s_ in side:
                        stream.wait_stream(s_)
        return g

    def timed(graph, replays, graph_rem):
        with torch.cuda.stream(stream):
            graph.replay()                                               # graph warm-up (uploads the exec graph)
            if graph_rem:
                graph_rem.replay()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(replays):
                graph.replay()
            if graph_rem:
                graph_rem.replay()
            ev1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        return ev0.elapsed_time(ev1), t0, time.perf_counter()

    graph = capture(g_steps)
    launches_per_step = (sum(h.launch_count() for h in batches) - launches_before) / g_steps
    replays, rem = K // g_steps, K % g_steps
    graph_rem = capture(rem) if rem else None

    # ---- timed region: exactly K steps, one stream, CUDA events on the launching stream, barrier + sync both sides;
    #      repeated `--repeats` times, the median repeat is reported
    runs = sorted((timed(graph, replays, graph_rem) for _ in range(max(1, args.repeats))), key=lambda r: r[0])
    ms_total, t_wall0, t_wall1 = runs[len(runs) // 2]
    ms_all = [r[0] for r in runs]

    # ---- same K steps with independent batches overlapped on 3 parallel graph branches (extra figure)
    n_ov = min(3, n_batches)
    graph_ov = capture(g_steps, n_ov)
    graph_ov_rem = capture(rem, n_ov) if rem else None
    ms_overlap = sorted(timed(graph_ov, replays, graph_ov_rem)[0] for _ in range(max(1, args.repeats)))[max(1, args.repeats) // 2]

    # ---- K-step rollout kernel: 64 steps per launch, uniform random policy drawn on the device
    roll_T = 64
    for h in batches:
        h.rollout(roll_T, None, policy_seed=1, **step_kw)
    torch.cuda.synchronize(dev)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_roll = max(2 * n_batches, 8)
    r0.record()
    for i in range(n_roll):
        batches[i % n_batches].rollout(roll_T, None, policy_seed=2 + i, **step_kw)
    r1.record()
    torch.cuda.synchronize(dev)
    roll_ms = r0.elapsed_time(r1)

    # ---- closed-loop rollout: integer linear policy evaluated on the device from each step's observation
    roll_policy_ms = None
    if batches[0].obs_dim > 0 and int(n_act.max().item()) <= 16:
        A = int(n_act.max().item())
        gw = torch.Generator(device=dev)
        gw.manual_seed(7)
        w_pol = torch.randint(-9, 10, (batches[0].obs_dim, A), generator=gw, device=dev, dtype=torch.int32)
        b_pol = torch.randint(-30, 31, (A,), generator=gw, device=dev, dtype=torch.int32)
        for h in batches:
            h.rollout(roll_T, policy=(w_pol, b_pol), **step_kw)
        torch.cuda.synchronize(dev)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for i in range(n_roll):
            batches[i % n_batches].rollout(roll_T, policy=(w_pol, b_pol), **step_kw)
        q1.record()
        torch.cuda.synchronize(dev)
        roll_policy_ms = q0.elapsed_time(q1)

    # ---- eager (one python call per launch) figure, for the launch-bound picture
    n_eager = min(K, 2000)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_eager):
        one_step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    eager_ms = e0.elapsed_time(e1) / n_eager

    # ---- end-to-end through the host-buffer C-ABI call (ngw_step_host), pinned buffers, every step H2D + D2H
    n_e2e = min(K, 200 if envs <= 65536 else 40)
    host_acts = [a.cpu().numpy() for a in act_sets]
    for i in range(max(3, n_batches)):                                  # every handle allocates its pinned buffers here
        batches[i % n_batches].step_host(host_acts[i % n_sets], **step_kw)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    # depth-2 software pipeline over the rotating batches: batch i+1 is enqueued (H2D, launch, D2H on its own stream)
    # before the host waits for batch i, so the PCIe link never idles; every step still copies its inputs in and its
    # complete results out, and the results are read on the host
    batches[0].step_host_begin(host_acts[0], **step_kw)
    for i in range(n_e2e):
        if i + 1 < n_e2e:
            batches[(i + 1) % n_batches].step_host_begin(host_acts[(i + 1) % n_sets], **step_kw)
        obs, rew, dn, cost, res = batches[i % n_batches].step_host_end()
        checksum += float(rew[0]) + float(obs[0, 0])                    # touch the results on the host
    t_e2e = time.perf_counter() - t0
    # the plain blocking call, one batch at a time
    t0 = time.perf_counter()
    for i in range(n_e2e):
        obs, rew, dn, cost, res = batches[i % n_batches].step_host(host_acts[i % n_sets], **step_kw)
        checksum += float(rew[0]) + float(obs[0, 0])
    t_e2e_blocking = time.perf_counter() - t0
    t_timed_end = time.perf_counter()

    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary(t_wall0, t_timed_end)

    stats = torch.zeros(8, dtype=torch.float64, device=dev)
    for h in batches:
        stats += h.stats()
    times = torch.tensor([ms_total, t_e2e * 1e3, ms_overlap], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)                     # max over ranks
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)                     # episode statistics over NCCL
    ms_total, e2e_ms_total, ms_overlap = [float(x) for x in times.cpu().numpy()]

    if rank == 0:
        ms_per_step = ms_total / K
        value = world * envs * K / (ms_total * 1e-3)
        peak, peak_src = measured_hbm_peak()
        achieved = envs * bytes_step / (ms_per_step * 1e-3) / 1e9
        ov_achieved = envs * bytes_step / (ms_overlap / K * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')) as f:
                traffic = json.load(f).get(args.workload, {}).get('dram_bytes_per_launch')
        except Exception:
            pass
        d_obs = batches[0].obs_dim
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "envs_per_batch": envs, "batches_rotated": n_batches,
                       "l2": "inputs larger than L2: %d rotating batches, %.0f MB combined working set vs 126 MB L2"
                             % (n_batches, n_batches * per_batch_ws / 1e6),
                       "launch": "one stream, CUDA-graph replay of %d-step graphs (one kernel launch per step); eager "
                                 "python loop = %.2f us/step" % (g_steps, eager_ms * 1e3),
                       "per_gpu_envs": n_batches * envs, "parallelism": "independent shards x%d" % world,
                       "reset_error_flags": flags},
            "clocks": clocks,
            "e2e": {"value": world * envs * n_e2e / (e2e_ms_total * 1e-3), "unit": "env-steps/s",
                    "h2d_bytes_per_step": 4 * envs,
                    "d2h_bytes_per_step": (4 * d_obs + 4 + 4 + 1 + 1) * envs,
                    "steps": n_e2e,
                    "api": "ngw_step_host_begin/_end, pinned host buffers: H2D actions, one launch, D2H obs/reward/done/"
                           "step_cost/result per step; batch i+1 enqueued before waiting for batch i",
                    "blocking_value": world * envs * n_e2e / t_e2e_blocking},
            "gpu_launches": int(round(launches_per_step * K)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "ngw::step_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": bytes_step,
                         "algorithmic_bytes_per_launch": envs * bytes_step,
                         "avg_launch_us": ms_per_step * 1e3},
            "overlapped": {"note": "same K steps, independent batches on %d parallel graph branches (launches overlap)"
                                   % n_ov, "value": world * envs * K / (ms_overlap * 1e-3), "unit": "env-steps/s",
                           "us_per_step": ms_overlap / K * 1e3, "achieved_gbs": ov_achieved,
                           "frac_of_hbm_peak": ov_achieved / peak},
            "rollout": {"note": "ngw_rollout: %d steps per launch, on-device uniform random policy, tile resident in "
                                "shared memory; outputs are per-env sums + final observation" % roll_T,
                        "value": world * envs * roll_T * n_roll / (roll_ms * 1e-3), "unit": "env-steps/s"},
            "rollout_policy": None if roll_policy_ms is None else {
                "note": "ngw_rollout_policy: %d steps per launch, action = argmax(b + obs @ W) on the device from each "
                        "step's lidar observation" % roll_T,
                "value": world * envs * roll_T * n_roll / (roll_policy_ms * 1e-3), "unit": "env-steps/s"},
            "eager": {"value": envs / (eager_ms * 1e-3), "unit": "env-steps/s", "us_per_step": eager_ms * 1e3},
            "episode_stats": dict(zip(('steps', 'episodes', 'successes', 'reward_sum', 'cost_sum', 'resets',
                                       'invalid'), [float(x) for x in stats.cpu().numpy()[:7]])),
            "wall_ms_timed_region": (t_wall1 - t_wall0) * 1e3,
            "repeats": {"n": len(ms_all), "ms_per_step_min_med_max": [ms_all[0] / K, ms_all[len(ms_all) // 2] / K,
                                                                       ms_all[-1] / K]},
        }
        if world == 1 and not args.no_cpu_baseline and args.workload == 'C2':
            threads = os.cpu_count() or 1
            steps_cpu, dt = cpu_port_run(compiled[0], ENVS_PER_BATCH, 10 ** 9, 16, threads, time_budget_s=3.0)
            line["cpu_baseline"] = {
                "value": steps_cpu * ENVS_PER_BATCH / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
                "sample": "%d steps x %d envs (%.1f s wall, ~%.0f s of CPU work), C port of the reference path "
                          "(oracle/ngw_oracle.c), %d threads, no per-step barrier"
                          % (steps_cpu, ENVS_PER_BATCH, dt, dt * threads, threads)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=32768)
    ap.add_argument('--warmup', type=int, default=64)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--repeats', type=int, default=3, help='timed region is run this many times; the median is reported')
    ap.add_argument('--workload', default='C2', choices=['C2', 'C3', 'C4', 'C4-blocked', 'C5'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
