"""Round-2 golden vectors, generated like make_golden.py: the REFERENCE'S OWN SOURCES (imported from /root/reference,
unmodified) executed under the Taichi-semantics shim oracle/tishim.  Authoring container only; the .npz files are committed.

    python tests/golden/make_golden_r2.py

Writes
  normaldist.npz  erfinv / normaldist (ptina/common.py:337-352) on fixed inputs -- the MLT small-step kernel (mltpath.py:64)
  preview.npz     PreviewEngine().render() x 2 (engine/preview.py:19-41) on the 12x12 config-2 scene and on the textured mini
                  matball: film passes 1 (albedo) and 2 (normal), get_image of both
  tile.npz        PathEngine.render_tile / render_final.  That code sits in engine/path.py:95-128 inside a ''' string (disabled
                  in the reference; exams/benchtiles.py still calls it).  The generator cuts the text out of the reference file,
                  dedents it and attaches the three methods to the reference's PathEngine class -- nothing is retyped -- then runs
                  render_tile(0, 0, 1) on a 12x10 film.  The text's bounds checks are `x > nx` / `y > ny` (sic): x == nx writes
                  past the image and y == ny lands in pixel (x+1, 0).  Both effects are recorded separately (`film_text` is what
                  the text produces; `film` has the two out-of-range rows removed, which is what a bounds-correct implementation
                  renders) so the tests can hold the oracle and the CUDA path to the in-range part bit for bit / to tolerance.
"""
import os
import sys
import textwrap
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the shim and imports the reference package)

ti, common, things, enginepath, sobolmod = mg.ti, mg.common, mg.things, mg.enginepath, mg.sobolmod
FilmTable, BVHTree, Camera, ModelPool, Stack = mg.FilmTable, mg.BVHTree, mg.Camera, mg.ModelPool, mg.Stack
scenes = mg.scenes


def films(nx, ny):
    root = FilmTable().root.arr
    return [root[k, :nx * ny].copy().reshape(nx, ny, 4) for k in range(3)]


def attach_tile_methods():
    """engine/path.py:95-128: the ''' block after do_render, verbatim, compiled into methods of the reference's PathEngine."""
    src = open(os.path.join(mg.REF, 'ptina', 'engine', 'path.py')).read()
    a = src.index("'''", src.index('def do_render'))
    b = src.index("'''", a + 3)
    block = textwrap.dedent(src[a + 3:b])
    assert 'def render_tile' in block and 'def _render_tile' in block and 'def render_final' in block
    ns = dict(enginepath.__dict__)
    exec(compile(block, 'ptina/engine/path.py[95:128]', 'exec'), ns)
    for name in ('render_tile', '_render_tile', 'render_final'):
        setattr(enginepath.PathEngine, name, ns[name])
    return block


def main():
    t00 = time.time()
    mg.quiet(things.init_things)
    mg.quiet(enginepath.PathEngine)
    import ptina.engine.preview as enginepreview
    enginepreview.__dict__.update(ti.KERNEL_BUILTINS)
    mg.quiet(enginepreview.PreviewEngine)
    sob = sobolmod.SobolSampler()
    print(f'SobolSampler ready (time={int(sob.time[None])}) in {time.time() - t00:.1f}s', flush=True)

    # ---- S6: erfinv / normaldist --------------------------------------------------------------------------------------------
    rng = np.random.default_rng(77)
    samp = np.concatenate([rng.random(500), [0.5, 0.25, 0.75, 1e-6, 1 - 1e-6, 0.001, 0.999, 0.4999999, 0.5000001]]).astype(np.float32)
    nd = np.array([float(common.normaldist(ti.f32(float(s)))) for s in samp], np.float32)
    xs = (samp * np.float32(2) - np.float32(1)).astype(np.float32)
    ei = np.array([float(common.erfinv(ti.f32(float(x)))) for x in xs], np.float32)
    np.savez_compressed(os.path.join(HERE, 'normaldist.npz'), samp=samp, normaldist=nd, erfinv_x=xs, erfinv=ei)
    print('normaldist done', flush=True)

    # ---- f1: PreviewEngine ----------------------------------------------------------------------------------------------------
    out = {}
    for name, sc in (('cornell_monkey', dict(scenes.cornell_monkey(), size=(12, 12))), ('mini_matball', scenes.mini_matball())):
        t0 = time.time()
        scenes.apply(mg.RefWorker(), sc)
        FilmTable().clear()
        ks = []
        for _ in range(2):
            mg.quiet(enginepreview.PreviewEngine().render)
            ks.append(int(sob.time[None]))
        nx, ny = sc['size']
        f = films(nx, ny)
        out.update({f'{name}_ks': np.array(ks), f'{name}_film1': f[1], f'{name}_film2': f[2], f'{name}_film0': f[0],
                    f'{name}_image1': np.asarray(FilmTable().get_image(1), np.float32), f'{name}_image2': np.asarray(FilmTable().get_image(2), np.float32)})
        out.update({f'{name}_{k}': v for k, v in mg.scene_arrays(sc).items()})
        print(f'preview {name}: frames {ks} in {time.time() - t0:.1f}s', flush=True)
    np.savez_compressed(os.path.join(HERE, 'preview.npz'), **out)

    # ---- f4: render_tile from the reference's disabled text ---------------------------------------------------------------------
    t0 = time.time()
    attach_tile_methods()
    sc = dict(scenes.cornell_monkey(), size=(12, 10))
    scenes.apply(mg.RefWorker(), sc)
    FilmTable().clear()
    nx, ny = sc['size']
    k_before = int(sob.time[None])
    mg.quiet(enginepath.PathEngine().render_tile, 0, 0, 1)           # samples = 1 -> m in {0, 1} (`m > samples: continue`)
    k_tile = int(sob.time[None])
    root = FilmTable().root.arr[0].copy()
    text = root[:nx * ny].reshape(nx, ny, 4).copy()
    # what lands outside 0 <= x < nx, 0 <= y < ny under the text's `>` checks: (x = nx, y) -> index nx*ny + y (past the image);
    # (x, y = ny) -> index x*ny + ny = pixel (x+1, 0) for x < nx-1.  Re-render exactly those contributions to subtract them.
    spill = root[nx * ny:nx * ny + ny + 1].copy()
    eng = enginepath.PathEngine()
    alias = np.zeros((nx, ny, 4), np.float32)
    Stack().set(0)
    for x in range(nx):
        y = ny
        for m in range(2):
            rg = sobolmod.SobolSampler().get_proxy(mg.sampling.wanghash3(ti.i32(x), ti.i32(y), ti.i32(m)))
            dx, dy = common.random2(rg)
            fx = (ti.i32(x) + dx) / FilmTable().nx * 2 - 1
            fy = (ti.i32(y) + dy) / FilmTable().ny * 2 - 1
            clr = enginepath.path_trace(Camera().generate(fx, fy), rg)
            if x + 1 < nx:
                alias[x + 1, 0] += np.array(mg.fl(clr) + [1.0], np.float32)
    Stack().unset()
    np.savez_compressed(os.path.join(HERE, 'tile.npz'), **mg.scene_arrays(sc), k_before=np.int32(k_before), k_tile=np.int32(k_tile), samples=np.int32(1),
                        film_text=text, alias=alias, spill=spill)
    print(f'tile: k {k_before} -> {k_tile}, weights {np.unique(text[..., 3]).tolist()} in {time.time() - t0:.1f}s', flush=True)
    print('round-2 golden vectors written in', time.time() - t00, 's')


if __name__ == '__main__':
    main()
