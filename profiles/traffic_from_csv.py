"""ncu --csv launch lists (dram__bytes_read.sum, dram__bytes_write.sum per launch) -> profiles/step_kernel_traffic.json.
    python profiles/traffic_from_csv.py <32-batch csv> <7-batch csv>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def per_launch(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0] != '' and not r[0].startswith('==')]
    hdr = rows[0]
    i_id, i_name, i_unit, i_val = hdr.index('ID'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    acc = {}
    for r in rows[1:]:
        if r[i_name] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            acc.setdefault(r[i_id], {})[r[i_name]] = float(r[i_val].replace(',', '')) * scale.get(r[i_unit], 1.0)
    n = len(acc)
    rd = sum(v.get('dram__bytes_read.sum', 0.0) for v in acc.values()) / max(n, 1)
    wr = sum(v.get('dram__bytes_write.sum', 0.0) for v in acc.values()) / max(n, 1)
    return n, rd, wr


def main():
    big, small = sys.argv[1], sys.argv[2]
    n, rd, wr = per_launch(big)
    n7, rd7, wr7 = per_launch(small)
    alg = 65536 * 446
    out = {"C2": {
        "dram_bytes_per_launch": int(rd + wr), "dram_read_bytes_per_launch": int(rd), "dram_write_bytes_per_launch": int(wr),
        "launches_averaged": n, "rotating_batches": 32, "working_set_mb": 935, "algorithmic_bytes_per_launch": alg,
        "kernel": "ngw::step1w_kernel<1, 8>",
        "source": "profiles/%s: ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum "
                  "-k regex:step1w_kernel -s 64 -c 64 python profiles/prof_step.py C2 160 32 (two full rotations of 32 "
                  "batches, write-back included)" % os.path.basename(big),
        "other_rotations": {"7_batches_205_MB": {
            "dram_bytes_per_launch": int(rd7 + wr7), "read": int(rd7), "write": int(wr7), "launches_averaged": n7,
            "source": "profiles/%s (-s 14 -c 28, four rotations)" % os.path.basename(small)}},
        "note": "ratio to the algorithmic bytes: %.2f with 32 rotating batches (7.4 x L2), %.2f with 7 (1.6 x L2, the L2 keeps "
                "part of the rotation); bench.py rotates 18 batches (4.2 x L2) and reports both as roofline.l2_sensitivity"
                % ((rd + wr) / alg, (rd7 + wr7) / alg)}}
    with open(os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["C2"])[:400])


if __name__ == '__main__':
    main()
