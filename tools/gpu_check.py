"""First-light diagnostics on a GPU box: parity vs the oracle with verbose numbers (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from ptina_b200 import scenes, worker, _native
from ptina_b200.engine import PathEngine, BruteEngine

worker.init()
ctx = _native.context()
print('L2 GB/s (64MB):', ctx.measure_l2(64, 20), ' (32MB):', ctx.measure_l2(32, 20))
for name in sys.argv[1:] or ['cornell_boxes', 'cornell_monkey', 'matball', 'mega_small']:
    sc = scenes.CONFIGS[name]()
    nx, ny = sc['size']
    if name != 'mega_small':
        sc['size'] = (128, 128 * ny // nx)
    ctx.sobol_reset()
    t = time.time(); scenes.apply(worker, sc); t_gpu = time.time() - t
    ref = oracle.Oracle(); scenes.apply(ref, sc)
    info = ctx.tree
    print(f'== {name}: n={info.n} sweeps={info.aabb_sweeps} valid={info.valid} depth={info.depth} policy={info.policy} build_ms={info.build_ms:.3f} (oracle valid depth {ref.validate_tree()})')
    a, b = ctx.export_tree(), ref.export_tree()
    for key in ('mc', 'id', 'child', 'leaf', 'bmin', 'bmax'):
        print(f'   tree {key}: equal={np.array_equal(a[key], b[key])}')
    k = 65
    pa, pb = ctx.trace_primary(k), ref.primary(k)
    print('   rays bit-equal:', np.array_equal(pa['rays'].view(np.int32), pb['rays'].view(np.int32)),
          ' hit ids equal:', np.array_equal(pa['index'], pb['index']), ' mismatches:', int((pa['index'] != pb['index']).sum()),
          ' depth bit-equal:', np.array_equal(pa['depth'].view(np.int32), pb['depth'].view(np.int32)),
          ' uv bit-equal:', np.array_equal(pa['uv'].view(np.int32), pb['uv'].view(np.int32)))
    ctx.set_traversal(_native.TRAVERSE_REFERENCE)
    pc = ctx.trace_primary(k)
    ctx.set_traversal(_native.TRAVERSE_AUTO)
    print('   literal-policy hit ids equal:', np.array_equal(pc['index'], pb['index']), ' depth:', np.array_equal(pc['depth'].view(np.int32), pb['depth'].view(np.int32)))
    eng = _native.ENGINE_BRUTE if sc['engine'] == 'brute' else _native.ENGINE_PATH
    oeng = oracle.ENGINE_BRUTE if sc['engine'] == 'brute' else oracle.ENGINE_PATH
    sa, sb = ctx.render_sample(eng, k), ref.render_sample(oeng, k)
    d = np.abs(sa - sb)
    rel = d / np.maximum(np.abs(sb), 1e-3)
    print(f'   sample radiance: max abs {d.max():.3e}  max rel {rel.max():.3e}  pixels rel>1e-4: {(rel.max(axis=2) > 1e-4).sum()} / {sa.shape[0]*sa.shape[1]}  nan gpu {np.isnan(sa).sum()} ref {np.isnan(sb).sum()}')
    worker.clear(); ref.clear()
    ctx.set_counting(True, True); ctx.reset_counters()
    spp = 8
    t = time.time(); ctx.render(eng, spp); ctx.synchronize(); dt = time.time() - t
    cnt = ctx.counters(); ms = ctx.stage_ms()
    ctx.set_counting(False, False)
    t = time.time(); ocnt = ref.render(oeng, spp); odt = time.time() - t
    ia, ib = worker.get_image()[..., :3], ref.get_image()[..., :3]
    rmse = np.sqrt(((ia - ib) ** 2).mean()) / max(ib.mean(), 1e-6)
    print(f'   {spp} spp image rel-RMSE vs oracle: {rmse:.3e}; gpu rays {cnt["rays"]} oracle rays {ocnt["rays"]}; gpu {dt*1e3:.1f} ms stage {ms}; oracle {odt:.2f} s ({ocnt["rays"]/odt/1e6:.2f} Mrays/s)')
    print(f'   counters gpu {cnt}')
    print(f'   counters ref {ocnt}')
