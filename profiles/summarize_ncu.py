"""ncu report -> small JSON summary kept under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).
    python profiles/summarize_ncu.py gpurun_out/x.ncu-rep profiles/x.json ["note"]
Per captured launch: duration, DRAM bytes, instructions, occupancy limits; aggregated over the first launch: warp-stall
samples by reason and the source lines that collect the most samples / instructions (needs -lineinfo + --import-source)."""
import csv
import io
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'launch__occupancy_limit_barriers', 'launch__waves_per_multiprocessor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__inst_executed.sum.per_cycle_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']


def ncu(rep, *args):
    return subprocess.run(['ncu', '-i', rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    rows = list(csv.reader(io.StringIO(ncu(rep, '--page', 'raw', '--csv'))))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {'kernel': d.get('Kernel Name', '')[:90]}
        for k in WANT:
            if d.get(k) not in (None, ''):
                try:
                    rec[k + ' [' + units[hdr.index(k)] + ']'] = float(d[k])
                except ValueError:
                    rec[k] = d[k]
        launches.append(rec)
    src = list(csv.reader(io.StringIO(ncu(rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'))))
    cur, h, agg, kern = None, None, {}, 0
    for r in src:
        if not r:
            continue
        if r[0] == 'Kernel Name':
            kern += 1
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            h = r
            continue
        if r[0] not in ('', '...') and h:
            d = dict(zip(h, r))
            try:
                key = (cur, int(r[0]), r[1].strip()[:110])
                v = agg.setdefault(key, [0, 0, {}])
                v[0] += int(d['# Samples'])
                v[1] += int(d['Instructions Executed'])
                for name in h:
                    if name.startswith('stall_') and 'Not Issued' not in name and d.get(name):
                        x = int(d[name])
                        if x:
                            v[2][name] = v[2].get(name, 0) + x
            except (ValueError, KeyError):
                pass
    n_l = max(len(launches), 1)
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    stalls = {}
    for v in agg.values():
        for k, x in v[2].items():
            stalls[k] = stalls.get(k, 0) + x
    def line(k, v):
        return {'where': '%s:%d' % (k[0], k[1]), 'source': k[2], 'samples_pct': round(100.0 * v[0] / tot_s, 1),
                'warp_instructions_per_launch': v[1] // n_l, 'top_stalls': dict(sorted(v[2].items(), key=lambda kv: -kv[1])[:3])}
    summary = {'report': rep.split('/')[-1], 'note': note, 'launches': launches,
               'warp_stall_samples': dict(sorted(stalls.items(), key=lambda kv: -kv[1])),
               'warp_instructions_per_launch': tot_i // n_l,
               'hot_lines_by_samples': [line(k, v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]],
               'hot_lines_by_instructions': [line(k, v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]]}
    with open(out, 'w') as f:
        json.dump(summary, f, indent=1)
    print('wrote', out, len(launches), 'launches')


if __name__ == '__main__':
    main()
