"""Per-source-line instruction counts of one profiled launch: joins the SASS page of an `ncu --set full --import-source on` report
with the line table of the cubin (nvdisasm -g), by instruction offset.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep LAUNCH_INDEX ptina_b200/build [top]     # or ptina_b200/libptina_b200.so

Needs the ncu CLI, cuobjdump and nvdisasm; no GPU."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_page(rep, launch):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(launch), '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = next(r[1] for r in rows if r and r[0] == 'Kernel Name')
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    ix = {k: i for i, k in enumerate(rows[h])}
    body = []
    for r in rows[h + 1:]:
        if not r or not r[0].startswith('0x'):
            break
        body.append(r)
    base = int(body[0][0], 16)
    recs = []
    for r in body:
        recs.append({'off': int(r[0], 16) - base, 'sass': r[ix['Source']].strip(), 'inst': int(r[ix['Instructions Executed']]), 'thr': int(r[ix['Thread Instructions Executed']]),
                     'samples': int(r[ix['# Samples']] or 0)})
    return name, recs


def line_table(so, mangled_pat):
    # `so`: the library, or a directory of object files (the two builds of shade.cu carry the same cubin name inside the library)
    bins = [os.path.join(so, f) for f in sorted(os.listdir(so)) if f.endswith('.o')] if os.path.isdir(so) else [so]
    cubins = []
    for b in bins:
        tmp = tempfile.mkdtemp()
        subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(b)], cwd=tmp, capture_output=True)
        cubins += [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')]
    tabs = {}
    for f in cubins:
        txt = subprocess.run(['nvdisasm', '-g', '-c', f], capture_output=True, text=True).stdout
        fn, cur = None, None
        for ln in txt.splitlines():
            m = re.match(r'\s*\.section\s+\.text\.(\S+),', ln)
            if m:
                fn = m.group(1) + '@' + os.path.basename(os.path.dirname(f)); tabs[fn] = {}; cur = None
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
            if m and fn:
                tabs[fn][int(m.group(1), 16)] = cur
    return tabs


def demangle(s):
    return subprocess.run(['cu++filt', s], capture_output=True, text=True).stdout.strip()


def main():
    rep, launch, so = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    name, recs = sass_page(rep, launch)
    tabs = line_table(so, None)
    # the cubin function with the same base name, the same first template argument and the same number of instructions
    base = re.match(r'(?:void )?(?:<unnamed>::)?(\w+)', name).group(1)
    targ = re.search(r'<(?:\(\w+\))?(\w+)', name.replace('<unnamed>::', ''))
    cands = [fn for fn in tabs if base in fn and len(tabs[fn]) == len(recs) and (not targ or targ.group(1) in fn or targ.group(1).isdigit())]
    if not cands:
        sys.exit(f'no cubin function matches {name} ({len(recs)} instructions)')
    best = cands[0]
    print(f'# cubin function: {best}' + (f'  ({len(cands)} candidates)' if len(cands) > 1 else ''))
    tab = tabs[best]
    per = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in recs:
        key = tab.get(r['off'])
        for k, v in enumerate((r['inst'], r['thr'], r['samples'])):
            per[key][k] += v; tot[k] += v
    print(f'# {name}\n# {tot[0]} warp instructions, {tot[1] / max(tot[0], 1):.2f} lanes, {tot[2]} stall samples')
    print('| file:line | warp inst % | lanes | stall samples % |')
    for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'| {key[0]}:{key[1]} | {100.0 * v[0] / tot[0]:.2f} | {v[1] / max(v[0], 1):.1f} | {100.0 * v[2] / max(tot[2], 1):.2f} |' if key else f'| ? | {100.0 * v[0] / tot[0]:.2f} | - | - |')


if __name__ == '__main__':
    main()
