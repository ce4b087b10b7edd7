"""Warp / thread instruction counts of ONE bench step from an ncu metrics CSV -> profiles/inst_counts.json (what bench.py's
issue-slot roofline divides by the live step time).

On the GPU box (after the same command has exited 0 without ncu):
    ncu --profile-from-start off --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/inst_<scene>.csv python bench.py --one-step --scene <scene>
Here:
    python tools/ncu_inst_counts.py gpurun_out/inst_cornell_monkey.csv cornell_monkey 32 [more csv/scene/spp triples]
The workload is deterministic (fixed Sobol indices, fixed scene), so the counts are a property of the build, not of the run.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if r]
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    ix = {n: i for i, n in enumerate(rows[h])}
    per = {}
    for r in rows[h + 1:]:
        if len(r) <= ix['Metric Value']:
            continue
        name = r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        k = per.setdefault(name, {'launches': set(), 'warp_inst': 0.0, 'thread_inst': 0.0, 'us': 0.0, 'dram_bytes': 0.0})
        k['launches'].add(r[ix['ID']])
        v = float(r[ix['Metric Value']].replace(',', ''))
        m, unit = r[ix['Metric Name']], r[ix['Metric Unit']]
        if m == 'smsp__inst_executed.sum':
            k['warp_inst'] += v
        elif m == 'smsp__thread_inst_executed.sum':
            k['thread_inst'] += v
        elif m == 'gpu__time_duration.sum':
            k['us'] += v / 1e3 if unit == 'ns' else v * (1e3 if unit == 'ms' else 1.0)
        elif m.startswith('dram__bytes'):
            k['dram_bytes'] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    return per


def main():
    out_path = os.path.join(ROOT, 'profiles', 'inst_counts.json')
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    args = sys.argv[1:]
    for i in range(0, len(args), 3):
        path, scene, spp = args[i], args[i + 1], int(args[i + 2])
        per = parse(path)
        tot_w = sum(k['warp_inst'] for k in per.values()); tot_t = sum(k['thread_inst'] for k in per.values())
        tot_us = sum(k['us'] for k in per.values())
        out[scene] = {
            'spp': spp, 'warp_inst_per_step': tot_w, 'thread_inst_per_step': tot_t, 'lanes_per_inst': tot_t / max(tot_w, 1),
            'dram_bytes_per_step': sum(k['dram_bytes'] for k in per.values()), 'ncu_us_per_step_serialised': tot_us,
            'source': f'ncu smsp__inst_executed.sum / smsp__thread_inst_executed.sum over every launch of one step ({os.path.basename(path)})',
            'per_kernel': {n: {'launches': len(k['launches']), 'warp_inst': k['warp_inst'], 'lanes_per_inst': k['thread_inst'] / max(k['warp_inst'], 1),
                               'share_of_ncu_time': k['us'] / max(tot_us, 1e-9), 'dram_bytes': k['dram_bytes']}
                           for n, k in sorted(per.items(), key=lambda kv: -kv[1]['warp_inst']) if k['warp_inst'] > 0.002 * tot_w}}
        print(scene, f'{tot_w / 1e6:.1f} M warp instructions, {tot_t / max(tot_w, 1):.2f} lanes, {tot_us / 1e3:.2f} ms under ncu')
    json.dump(out, open(out_path, 'w'), indent=1)


if __name__ == '__main__':
    main()
