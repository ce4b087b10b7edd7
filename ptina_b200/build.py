"""Builds libptina_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'libptina_b200.so')
SOURCES = ['api.cu', 'lbvh.cu', 'wavefront.cu', 'shade.cu', 'shade.cu:fast']      # shade.cu is compiled twice (strict / fast), see its header
HEADERS = ['ptb_internal.h', 'ptb_math.cuh', 'ptb_shade.cuh', 'ptb_traverse.cuh', 'ptb_trace_kernel.cuh', 'ptb_wavefront.cuh', '../../include/ptina_b200.h']

# -fmad=false / -prec-div / -prec-sqrt / -ftz=false: IEEE binary32 in source order -- what makes Morton codes, primary
# rays and the box/triangle predicates bit-identical to a strict CPU evaluation (DESIGN.md "Arithmetic contract").
NVCC_BASE = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-ftz=false',
             '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '--expt-relaxed-constexpr']
STRICT = ['-fmad=false', '-prec-div=true', '-prec-sqrt=true']
# the second compilation of shade.cu (PTB_MODE_FAST): FMA contraction, approximate division and square root in the arithmetic that
# owes the reference 1e-5 / an image gate, not bits
FAST = ['-fmad=true', '-prec-div=false', '-prec-sqrt=false', '-DPTB_SHADE_FAST=1']
NVCC_FLAGS = NVCC_BASE + STRICT


def flags_for(src, fast_flags=None):
    if src.endswith(':fast'):
        return NVCC_BASE + (list(fast_flags) if fast_flags else FAST)
    return NVCC_BASE + STRICT


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (needed to build libptina_b200.so)')


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    files = [os.path.join(CSRC, f.split(':')[0]) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in files)


def build(force=False, verbose=False, defines=(), out=None, fast_flags=None):
    """defines / out: tuning variants (`-DNAME=value` ... into another .so, selected at run time with PTINA_B200_LIB)."""
    if not force and not stale() and not defines and out is None and not fast_flags:
        return SO
    out = out or SO
    tag = '' if out == SO else '_' + os.path.basename(out).replace('.so', '')
    nvcc = _nvcc()
    env = dict(os.environ)
    if os.path.exists('/usr/bin/g++'):
        ccbin = ['-ccbin', '/usr/bin/g++']
    else:
        ccbin = []
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu:fast', '_fast.cu').replace('.cu', tag + '.o'))
        cmd = [nvcc] + ccbin + flags_for(src, fast_flags) + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src.split(':')[0]), '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(log)
        if p.returncode:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [nvcc] + ccbin + ['-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', out] + objs + ['-lcudart']
    subprocess.check_call(cmd, env=env)
    return out


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv if a.startswith('-D')]
    outs = [a[6:] for a in sys.argv if a.startswith('--out=')]
    fast = [f for a in sys.argv if a.startswith('--fast-flags=') for f in a[13:].split(',')]      # tuning: other flags for the fast build of shade.cu
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, defines=defs, out=os.path.abspath(outs[0]) if outs else None,
                fast_flags=(fast + ['-DPTB_SHADE_FAST=1']) if fast else None))
