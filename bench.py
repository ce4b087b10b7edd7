"""bench.py -- headline benchmark: PathEngine on the Cornell+monkey scene (BASELINE.json configs[1]:
978 triangles, 512x512, 32 spp, LBVH + MIS area light), metric Mrays/s (extend + shadow rays traced per second).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scene NAME]

A step = one full frame: `spp` calls of PathEngine.render() on one GPU (N GPUs: every rank renders `spp` samples of the
same frame at rank-interleaved Sobol indices -> weak scaling, film combined with one NCCL reduce per step).
`value` times steps with the scene resident in HBM; `e2e` times the whole user-visible call sequence from pinned host
buffers: load_model (H2D) + build_tree + render + get_image (D2H).
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

METRIC, UNIT = 'Mrays/s', 'Mrays/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ptina_b200')
    ap.add_argument('--scene', default='cornell_monkey')
    ap.add_argument('--spp', type=int, default=0)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


def workload(args):
    from ptina_b200 import scenes
    sc = scenes.CONFIGS[args.scene]()
    if args.spp:
        sc['spp'] = args.spp
    if args.scene == 'mega' and not args.spp:
        sc['spp'] = 4          # a bench step of config 4 is 4 spp at 1080p (1024 spp is the quality target, not a step)
    if args.scene == 'matball' and not args.spp:
        sc['spp'] = 16
    return sc


def config_of(sc, n_gpus):
    nx, ny = sc['size']
    return {'workload': f"{sc['name']}: {len(sc['mtlids'])} tris, {nx}x{ny}, {sc['spp']} spp/step/GPU, engine={sc['engine']}",
            'scene': sc['name'], 'tris': int(len(sc['mtlids'])), 'resolution': [nx, ny], 'spp_per_step_per_gpu': sc['spp'],
            'parallelism': f'sample-range x{n_gpus}' if n_gpus > 1 else 'single GPU',
            'l2_policy': 'path state per step (>=0.9 GB) exceeds the 126 MB L2, so every step starts cold'}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe): NVML in-process every few ms
    (nvidia_ml_py), falling back to polling nvidia-smi."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.max_mhz, self.source = index, [], False, None, 'nvidia-smi'
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            phys = int(vis.split(',')[index]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, 'nvml'
        except Exception:
            self.nvml = None

    def _nvml_sample(self):
        n = self.nvml
        mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        get = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.handle))
        bits = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}
        return [str(mhz), str(self.max_mhz)] + ['Active' if r & bits[k] else 'Not Active' for k in self.NAMES]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._nvml_sample())
                    time.sleep(0.004)
                    continue
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock sampling unavailable']}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2 + i].lower().startswith('active') for s in self.samples if len(s) > 2 + i)]
        return {'sm_mhz': mhz[len(mhz) // 2] if mhz else None, 'sm_max_mhz': int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                'reasons': reasons, 'samples': len(self.samples), 'source': self.source}


def cpu_reference(sc, seconds, threads=0, max_frames=1):
    """The oracle (CPU restatement of the reference algorithm) on all host threads, on a bounded sample: whole frames at
    the workload's resolution, as many samples per pixel as fit in ~`seconds`."""
    import oracle
    from ptina_b200 import scenes
    if threads <= 0:            # every core this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    ref = oracle.Oracle()
    scenes.apply(ref, sc)
    eng = oracle.ENGINE_BRUTE if sc['engine'] == 'brute' else oracle.ENGINE_PATH
    t0 = time.time(); cnt = ref.render(eng, 1, nthreads=threads); dt1 = time.time() - t0       # warm-up + calibration
    n = int(max(1, min(sc['spp'] * max_frames, seconds / max(dt1, 1e-3))))
    t0 = time.time(); cnt = ref.render(eng, n, nthreads=threads); dt = time.time() - t0
    nx, ny = sc['size']
    return {'value': cnt['rays'] / dt / 1e6, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': f"{n} spp of the {nx}x{ny} frame ({cnt['rays']} rays, {dt:.2f} s), OpenMP over pixel rows",
            'spp_per_s': n / dt, 'seconds': dt, 'rays': cnt['rays'],
            'note': 'C++ restatement of the Taichi algorithm (unordered DFS, per-pixel megakernel); Taichi is not installable in this image'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sc = workload(args)
    steps = max(1, args.steps)
    per_step = max(1.0, min(args.cpu_seconds, 150.0 / (steps + args.warmup)))
    vals = []
    for i in range(args.warmup + steps):
        r = cpu_reference(sc, per_step)
        if i >= args.warmup:
            vals.append(r)
    tot_rays = sum(v['rays'] for v in vals); tot_s = sum(v['seconds'] for v in vals)
    v = tot_rays / tot_s / 1e6
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
            'ms_per_step': tot_s / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_of(sc, args.gpus),
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': vals[-1]['cores'], 'kind': 'port', 'sample': vals[-1]['sample'] + f' per step x {steps} steps'},
            'spp_per_s': sum(v_['spp_per_s'] * v_['seconds'] for v_ in vals) / tot_s,
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from ptina_b200 import scenes, worker, _native
    from ptina_b200 import dist as pdist
    sc = workload(args)
    nx, ny = sc['size']
    spp = sc['spp']
    eng = {'path': _native.ENGINE_PATH, 'brute': _native.ENGINE_BRUTE}[sc['engine']]
    worker.init(device=local)
    ctx = _native.context()
    scenes.apply(worker, sc)
    info = ctx.tree

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def frame():
        # one step: this rank's share of the frame (rank-interleaved Sobol indices), then the film reduce
        k0 = ctx.sobol_time + 1
        first, count, stride = pdist.shard_range(k0, spp * world, rank, world)
        ctx.render_range(eng, first, count, stride)
        ctx.sobol_time = ctx.sobol_time + spp * world
        if world > 1:
            pdist.reduce_film_pass(0, 0)

    # ---- counters (untimed pass with the counting instantiation; the workload is deterministic) ----
    ctx.sobol_reset(); worker.clear()
    ctx.set_counting(True, False); ctx.reset_counters()
    frame(); ctx.synchronize()
    cnt = ctx.counters()
    ctx.set_counting(False, False)
    rays_per_step = cnt['rays']

    # ---- device-timed steps, inputs resident in HBM ----
    ctx.sobol_reset(); worker.clear()
    for _ in range(args.warmup):
        frame()
    ctx.set_counting(False, False); ctx.reset_counters()
    sampler = ClockSampler(local); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        frame()
    ev1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=10)      # an nvidia-smi query still in flight perturbs the host-synchronous e2e loop below
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches()
    # per-stage device times: a separate pass with CUDA events around every stage.  It runs the stages one after the other; the
    # timed region above overlaps the shadow stage of a bounce with the extend stage of the next one (two streams), where a
    # per-kernel duration is not defined.
    nprof = max(1, min(args.steps, 5))
    ctx.set_counting(False, True); ctx.reset_counters()
    for _ in range(nprof):
        frame()
    ctx.synchronize()
    stage = ctx.stage_ms()
    ctx.set_counting(False, False)
    t = torch.tensor([ms], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # every rank traces (statistically) the same number of rays per step; count this rank's exactly and sum over ranks
    r = torch.tensor([float(rays_per_step)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    total_rays_per_step = float(r.item())
    value = total_rays_per_step * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the public API with HOST buffers (pinned), H2D + build + render + D2H inside the timed region ----
    verts_pin = torch.from_numpy(np.ascontiguousarray(sc['vertices'], dtype=np.float32)).pin_memory()
    mtl_pin = torch.from_numpy(np.ascontiguousarray(sc['mtlids'], dtype=np.int32)).pin_memory()
    img_pin = torch.empty((nx, ny, 4), dtype=torch.float32).pin_memory()
    from ptina_b200.model import ModelPool
    from ptina_b200.tree import BVHTree

    e2e_parts = {}

    def e2e_step():
        t = [time.perf_counter()]
        ModelPool().load(verts_pin.numpy(), mtl_pin.numpy()); t.append(time.perf_counter())
        BVHTree().build(); t.append(time.perf_counter())
        worker.clear()
        frame(); t.append(time.perf_counter())
        if rank == 0:
            ctx.get_image(0, out=img_pin.numpy())
        else:
            ctx.synchronize()
        t.append(time.perf_counter())
        for name, a, b in zip(('load', 'build', 'launch', 'wait+readback'), t[:-1], t[1:]):
            e2e_parts.setdefault(name, []).append(round((b - a) * 1e3, 2))
    for _ in range(max(1, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dbg = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        e2e_step()
        dbg.append((time.perf_counter() - t1) * 1e3)
    e1.record()
    if os.environ.get('PTB_BENCH_DEBUG'):
        print('e2e per-step wall ms:', [round(x, 2) for x in dbg], {k: v[-args.steps:] for k, v in e2e_parts.items()}, file=sys.stderr)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    e2e_steps_ms = [round(x, 3) for x in dbg]
    t = torch.tensor([e2e_ms], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = total_rays_per_step * args.steps / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks['hbm_gbs'], 'measured (MEASURED_PEAKS.json)') if 'hbm_gbs' in peaks else (6650.0, 'fallback (B200_PROFILING.md)')
        l2_gbs = ctx.measure_l2(64, 20)
        # dominant stage = BVH traversal (extend + shadow; each stage launch = k_trace_pre + k_trace_tree): algorithmic bytes per
        # launch = 64 B per node visit + 64 B per triangle test + 48 B per ray (32 B ray in, 16 B hit out)  -- DESIGN.md "Roofline"
        n_trav_launches = 5 * nprof * (2 if eng == _native.ENGINE_PATH else 1)
        ncu = {}
        try:    # DRAM bytes / issue-slot utilisation of the traversal kernels from the committed ncu capture of this build
            ncu = json.load(open(os.path.join(ROOT, 'profiles', 'traversal_ncu.json'))).get(sc['name'], {})
        except Exception:
            pass
        trav_ms = stage['extend'] + stage['shadow']
        alg_bytes_step = 64.0 * cnt['node_visits'] + 64.0 * cnt['tri_tests'] + 48.0 * cnt['rays']
        achieved = alg_bytes_step * nprof / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else None
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': config_of(sc, world),
                'spp_per_s': spp * world * args.steps / (ms * 1e-3),
                'rays_per_step': total_rays_per_step,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(verts_pin.numel() * 4 + mtl_pin.numel() * 4),
                        'd2h_bytes_per_step': int(img_pin.numel() * 4), 'ms_per_step': e2e_ms / args.steps, 'ms_steps': e2e_steps_ms,
                        'path': 'ModelPool.load(pinned host) + BVHTree.build + FilmTable.clear + PathEngine.render_range + FilmTable.get_image(host)'},
                'gpu_launches': int(launches),
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': (achieved / hbm_peak) if achieved else None, 'traffic': ncu.get('dram_bytes_per_launch'),
                             'kernel': 'k_trace_pre + k_trace_tree (BVH traversal: extend and shadow stages)', 'ncu': ncu or None, 'launches': n_trav_launches, 'avg_launch_ms': trav_ms / n_trav_launches,
                             'algorithmic_bytes_per_launch': alg_bytes_step * nprof / n_trav_launches,
                             'timing': f'CUDA events around every stage, serial pass of {nprof} steps after the timed region', 'peak_source': peak_src,
                             'l2_peak_gbs_measured': l2_gbs, 'frac_of_l2': (achieved / l2_gbs) if achieved else None,
                             'note': 'gather workload served from shared memory / L1 / L2 (the scene is cache resident): HBM is the schema bound; '
                                     'what binds is instruction issue (ncu: issue-slot utilisation, lanes per instruction -- profiles/)'},
                'stage_ms_per_step': {k: v / nprof for k, v in stage.items()},
                'schedule': 'shadow stage of bounce b overlapped with the extend stage of bounce b+1 (2 streams); stage_ms_per_step from a serial pass',
                'counters_per_step_rank0': cnt,
                'tree': {'n': info.n, 'depth': info.depth, 'valid': info.valid, 'policy': info.policy, 'build_ms': info.build_ms},
                'clocks': sampler.summary()}
        if not args.no_cpu and world == 1:      # rank 0 at N=1 only (torchrun also pins OMP_NUM_THREADS=1)
            line['cpu_baseline'] = {k: v for k, v in cpu_reference(sc, args.cpu_seconds, max_frames=8).items() if k in ('value', 'unit', 'cores', 'kind', 'sample', 'spp_per_s', 'note')}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)
