python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python - <<'PY' > gpurun_out/r02_l2_sensitivity.jsonl 2>&1
import json, sys
sys.path.insert(0, 'profiles'); sys.path.insert(0, '.')
import sweep, bench
orig = bench.batches_to_exceed_l2
for nb in (7, 12, 18, 32):
    sweep.np.ceil  # noqa
    import numpy as np
    # sweep.measure computes n_b itself: override through the L2 size constant
    def measure_nb(nb=nb):
        import torch
        from gym_novel_gridworlds_b200.runtime import BatchHandle
        desc, compiled, envs, rule, kw = bench.build_workload('C2')
        batches = []
        for b in range(nb):
            h = BatchHandle(compiled, envs, seed=0, first_env_gid=b * envs); h.reset(); batches.append(h)
        g = torch.Generator(device='cuda'); g.manual_seed(1)
        acts = [torch.randint(0, 10, (envs,), generator=g, device='cuda', dtype=torch.int32) for _ in range(4)]
        for i in range(2 * nb): batches[i % nb].step(acts[i % 4])
        torch.cuda.synchronize()
        st = torch.cuda.Stream(); gr = torch.cuda.CUDAGraph()
        n_graph = nb * max(1, 64 // nb)
        with torch.cuda.stream(st):
            with torch.cuda.graph(gr, stream=st):
                for i in range(n_graph): batches[i % nb].step(acts[i % 4])
            gr.replay()
        torch.cuda.synchronize()
        res = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(st)
                for _ in range(100): gr.replay()
                e1.record(st)
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / (100 * n_graph) * 1e3)
        for h in batches: h.close()
        return sorted(res)[1]
    us = measure_nb()
    print(json.dumps({"batches": nb, "working_set_mb": nb * 65536 * 446 / 1e6, "us_per_step": round(us, 3), "frac": round(65536 * 446 / (us * 1e-6) / 1e9 / 6552.3, 4)}), flush=True)
PY
cat gpurun_out/r02_l2_sensitivity.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -c 300 gpurun_out/r02_bench_c.err
