"""Turns ncu output brought back from a GPU box into the small text summaries kept under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  profiles/r1_k_trace_full.md     # `ncu --set full` capture
  python tools/ncu_summary.py list gpurun_out/launches.csv  profiles/r1_launches.md         # `--metrics gpu__time_duration.sum` launch list

`rep` needs the ncu CLI (reads the report with --page raw / --page source); no GPU needed.
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__sass_average_branch_targets_threads_uniform.pct', 'sm__sass_inst_executed_op_local_ld.sum', 'sm__sass_inst_executed_op_local_st.sum',
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(rep, dst):
    rows = ncu_csv(rep, 'raw')
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f'# ncu --set full summary of `{rep}`', '',
             'Read with `ncu -i <rep> --page raw --csv`; one column per profiled launch.  Times under ncu are cold-cache and',
             'serialised (never bench values).', '']
    launches = rows[2:]
    names = [r[ix['Kernel Name']][:60] for r in launches]
    lines.append('| metric | unit | ' + ' | '.join(f'#{i}' for i in range(len(launches))) + ' |')
    lines.append('|---|---|' + '---|' * len(launches))
    lines.append('| kernel | | ' + ' | '.join(names) + ' |')
    for k in RAW_KEYS:
        if k in ix:
            lines.append(f'| {k} | {units[ix[k]]} | ' + ' | '.join(r[ix[k]] for r in launches) + ' |')
    # instruction mix of the first launch from the SASS page
    src = ncu_csv(rep, 'source', ['--launch-count', '1'])
    try:
        h = next(i for i, r in enumerate(src) if r and r[0] == 'Address')
        six = {name: i for i, name in enumerate(src[h])}
        mix, lanes, tot, thr = collections.Counter(), collections.Counter(), 0, 0
        seen = set()
        for r in src[h + 1:]:
            if len(r) <= six['Thread Instructions Executed'] or r[0] in seen:
                continue
            seen.add(r[0])
            try:
                n, t = int(r[six['Instructions Executed']]), int(r[six['Thread Instructions Executed']])
            except ValueError:
                continue
            op = [o for o in r[six['Source']].split() if not o.startswith('@')][0].split('.')[0]
            mix[op] += n; lanes[op] += t; tot += n; thr += t
        lines += ['', f'## SASS instruction mix, launch #0: {tot} warp instructions, {thr / max(tot, 1):.2f} active lanes per instruction', '',
                  '| opcode | % of warp instructions | lanes |', '|---|---|---|']
        for op, n in mix.most_common(24):
            lines.append(f'| {op} | {100.0 * n / tot:.2f} | {lanes[op] / max(n, 1):.1f} |')
    except StopIteration:
        pass
    open(dst, 'w').write('\n'.join(lines) + '\n')


def summarize_list(path, dst):
    rows = [r for r in csv.reader(open(path, errors='replace')) if r]
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    ix = {name: i for i, name in enumerate(rows[h])}
    agg, order = collections.OrderedDict(), []
    for r in rows[h + 1:]:
        if len(r) <= ix['Metric Value'] or r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = r[ix['Kernel Name']].split('(')[0]
        v = float(r[ix['Metric Value']].replace(',', ''))
        unit = r[ix['Metric Unit']]
        us = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    lines = [f'# launch list `{path}` (`ncu --metrics gpu__time_duration.sum --clock-control none`)', '',
             'Per-launch times under ncu are cold-cache and serialised: the SHARE of each kernel is what to compare with the', 'bench\'s stage times.', '',
             '| kernel | launches | total us | share % | avg us |', '|---|---|---|---|---|']
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'| `{name}` | {n} | {us:.1f} | {100 * us / tot:.1f} | {us / n:.1f} |')
    lines.append(f'| total | {sum(a[0] for a in agg.values())} | {tot:.1f} | 100 | |')
    open(dst, 'w').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    mode, src, dst = sys.argv[1:4]
    (summarize_rep if mode == 'rep' else summarize_list)(src, dst)
    print(open(dst).read())
