"""ncu --csv launch list (gpu__time_duration.sum per launch) -> per-kernel launch counts, total and average durations.
    python profiles/launches_from_csv.py profiles/r02_launches_bench_C2.csv profiles/r02_launches_bench_C2.json "<command>" """
import csv
import json
import sys


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ''
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith('==')]
    hdr = rows[0]
    i_k, i_m, i_u, i_v = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}
    acc = {}
    for r in rows[1:]:
        if r[i_m] != 'gpu__time_duration.sum':
            continue
        a = acc.setdefault(r[i_k][:72], [0, 0.0])
        a[0] += 1
        a[1] += float(r[i_v].replace(',', '')) * scale.get(r[i_u], 1.0)
    total = sum(v[1] for v in acc.values())
    out = {"command": cmd,
           "note": "launches of the bench process captured by ncu (set-up resets, warm-up, graph captures and the first timed "
                   "regions); per-launch times under ncu are serialised and cold-cache: shares, not absolutes",
           "kernels": [{"kernel": k, "launches": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2),
                        "share": round(v[1] / total, 3)} for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])]}
    json.dump(out, open(dst, 'w'), indent=1)
    for k in out["kernels"][:8]:
        print(k)


if __name__ == '__main__':
    main()
