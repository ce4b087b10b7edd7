"""SASS evidence for profiles/: opcode histogram, the memory / special instructions and a short excerpt around the node fetch of each hot
kernel, from the object files of the in-tree build (cuobjdump -sass; no GPU needed).

    python tools/sass_excerpts.py > profiles/r2_final_sass_excerpts.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, 'ptina_b200', 'build')
KERNELS = [   # (object, mangled-name regex, title, regex of the instruction the excerpt is centred on)
    ('wavefront.o', r'k_trace_treeI8ExtendIOLb0ELi768ELb1ELb0ELb0ELb1E', 'k_trace_tree<ExtendIO, resident 64-byte nodes, S16 stack> (configs 1-2)', r'LDS\.128'),
    ('wavefront.o', r'k_trace_treeI8ExtendIOLb0ELi768ELb1ELb1ELb0ELb1E', 'k_trace_tree<ExtendIO, resident quantised nodes, S16 stack> (config 3)', r'PRMT'),
    ('wavefront.o', r'k_trace_treeI8ExtendIOLb0ELi128ELb0ELb1ELb1ELb0E', 'k_trace_tree<ExtendIO, 4-wide quantised nodes out of global memory> (config 4)', r'LDG\.E\.[A-Z0-9.]*256'),
    ('wavefront.o', r'k_trace_treeI8ShadowIOLb0ELi768ELb1ELb0ELb0ELb1E', 'k_trace_tree<ShadowIO, resident> (configs 1-2)', r'RED'),
    ('wavefront.o', r'k_trace_preI8ExtendIOLb0E', 'k_trace_pre<ExtendIO>', r'FMNMX3'),
    ('shade.o', r'k_shadeILi0E', 'k_shade<PATH>, parity build', r'FCHK'),
    ('shade_fast.o', r'k_shadeILi0E', 'k_shade<PATH>, PTB_MODE_FAST build', r'MUFU\.RCP'),
]
SPECIAL = re.compile(r'^(LDG|STG|LDS|STS|LDL|STL|LDGSTS|LDGDEPBAR|ATOM|ATOMG|ATOMS|RED|REDG|FCHK|MUFU|FMNMX3|PRMT|BAR|VOTE|BRA\.DIV|UBLKCP|UTMALDG|HMMA|UTC|SHFL|MATCH|REDUX)')


def functions(obj):
    txt = subprocess.run(['cuobjdump', '-sass', os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    out, name = {}, None
    for ln in txt.splitlines():
        m = re.match(r'\s*Function : (\S+)', ln)
        if m:
            name = m.group(1); out[name] = []
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4})\*/\s+(.*?);', ln)
        if m and name:
            out[name].append((m.group(1), m.group(2).strip()))
    return out


def main():
    print('# SASS excerpts of the hot kernels (final round-2 build, `cuobjdump -sass ptina_b200/build/*.o`, sm_100a; `tools/sass_excerpts.py`)\n')
    print('What to look for: `LDG.E…256` = one-instruction 256-bit fetches of quantised nodes (two per 4-wide node) and of the streamed 64-byte triangle')
    print('record (`.NA` = no L1 allocation); `LDS.128` = resident nodes in shared memory; `PRMT` = decode of a 15-bit plane, its selector a per-ray')
    print('register (entry / exit plane by the ray\'s sign); `FMNMX3` = 3-input max / min of the slab distances; `LDGSTS.E.BYPASS.128` = `cp.async` staging')
    print('of the tree-queue tiles; `LDS` / `STS` (32-bit) = the S16 traversal stack; `REDG.E.ADD.F32.FTZ.RN` = the shadow contribution (DESIGN.md section 2,')
    print('exception i); `MUFU.RCP` + `FCHK` = IEEE division (parity build of `k_shade`), `FCHK` absent from the `PTB_MODE_FAST` build.  No `HMMA` / `UTC*MMA`:')
    print('no stage is a dense contraction.  No `UTMALDG` / `UBLKCP`: DESIGN.md section 9.\n')
    cache = {}
    for obj, pat, title, centre in KERNELS:
        fns = cache.setdefault(obj, functions(obj))
        names = [n for n in fns if re.search(pat, n)]
        if not names:
            print(f'## {title}\n\n(not found: {pat})\n'); continue
        name = names[0]
        ins = fns[name]
        ops = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', i).split()[0].split('.')[0] for _, i in ins)
        spec = collections.Counter()
        for _, i in ins:
            op = re.sub(r'^@!?U?P\d+\s+', '', i).split()[0]
            if SPECIAL.match(op):
                spec[op] += 1
        print(f'## {title}\n\n`{name}` — {len(ins)} instructions\n')
        print('opcode histogram (top 14): ' + ', '.join(f'{k} {v}' for k, v in ops.most_common(14)) + '\n')
        print('memory / special instructions: ' + ', '.join(f'`{k}` x{v}' for k, v in sorted(spec.items())) + '\n')
        idx = next((k for k, (_, i) in enumerate(ins) if re.search(centre, i)), None)
        if idx is not None:
            lo, hi = max(0, idx - 6), min(len(ins), idx + 14)
            print('```')
            for a, i in ins[lo:hi]:
                print(f'        /*{a}*/  {i} ;')
            print('```\n')


if __name__ == '__main__':
    main()
