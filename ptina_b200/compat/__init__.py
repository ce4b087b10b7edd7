"""Name compatibility for unmodified PTina driver scripts (exams/*.py, the Blender add-on's imports).

This directory holds two import-path packages:
  ptina/   the reference's module names (`ptina.things`, `ptina.engine.path`, `ptina.tools.readgltf`, `ptina.worker`, ...), each a
           re-export of the ptina_b200 module that implements it.  `from ptina.things import *` provides what the reference's does:
           `ti`, `np`, `init_things` and every singleton class.
  taichi/  a stand-in for the `taichi` module AS A DRIVER SCRIPT USES IT -- `ti.init(ti.cuda)`, the arch names, `ti.imshow`,
           `ti.imresize`, `ti.imwrite` -- nothing of the kernel language (there are no Taichi kernels here to compile).

Put the directory on the path to use them:

    python -m ptina_b200.compat exams/benchmark.py            # runs the script unmodified
    PYTHONPATH=$(python -m ptina_b200.compat --path) python exams/benchmark.py

They are deliberately NOT importable by default, so an installed Taichi / PTina is never shadowed by accident.
"""
import os
import runpy
import sys

PATH = os.path.dirname(os.path.abspath(__file__))


def install():
    """Make `import ptina` / `import taichi` resolve to the packages in this directory (idempotent)."""
    if PATH not in sys.path:
        sys.path.insert(0, PATH)
    return PATH


def run_script(path, argv=()):
    """Execute a driver script as __main__ with the compat packages on the path; returns its globals."""
    install()
    old = sys.argv
    sys.argv = [path, *argv]
    try:
        return runpy.run_path(path, run_name='__main__')
    finally:
        sys.argv = old
