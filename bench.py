"""bench.py -- headline benchmark: PathEngine on the Cornell+monkey scene (BASELINE.json configs[1]: 978 triangles, 512x512,
32 spp, LBVH + MIS area light), metric Mrays/s (extend + shadow rays traced per second), plus a `configs` block with every
BASELINE config at the run's GPU count.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scene NAME] [--quick]

Headline (`value`, `e2e`, `roofline`): a step = one full frame on every GPU -- rank g renders `spp` samples of the frame at
rank-interleaved Sobol indices (weak scaling), and for N > 1 the films are summed onto rank 0 with one NCCL reduce per step.
  value        steps timed on the device (CUDA events), scene resident in HBM
  e2e          the user-visible call sequence from pinned host buffers each step: ModelPool.load (H2D) + BVHTree.build +
               FilmTable.clear + render + FilmTable.get_image (D2H)
  e2e_percall  the same sequence with the reference's own call pattern -- 32 x PathEngine().render(), one sample per call
               (exams/benchmark.py:29-33); the library records the calls and submits them as one batch
  sustained    the same step repeated for >= 2 s (clocks under sustained load)
`configs`: config 1 (34 triangles), 2, 3 (BruteEngine 1024x1024, textures + environment), 4 (1.02 M triangles, 1920x1080: a fixed
  32-spp step SHARDED over the N GPUs + the 33 MB film reduce, and the full 1024-spp render), 5 (MLT, 2^18 chains sharded over the
  N GPUs), the strong-scaling split of a config-2 frame, and for N > 1 `film_check` (reduced film vs a 1-GPU render of the same
  Sobol indices) and what limits each sharded config (kernel tails vs reduce vs host enqueue), with numbers.
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
    ap.add_argument('--quick', action='store_true', help='headline only: no configs block, no sustained window')
    ap.add_argument('--one-step', action='store_true', help='profiling aid: set up, then exactly one step between cudaProfilerStart/Stop')
    return ap.parse_args()


STEP_SPP = {'cornell_boxes': 32, 'cornell_monkey': 32, 'matball': 32, 'mega': 32}


def workload(args, name=None):
    from ptina_b200 import scenes
    name = name or args.scene
    sc = scenes.CONFIGS[name]()
    sc['spp'] = args.spp if (args.spp and name == args.scene) else STEP_SPP.get(name, sc['spp'])
    return sc


def config_of(sc, n_gpus, mode='weak'):
    nx, ny = sc['size']
    per = f"{sc['spp']} spp/step/GPU" if mode == 'weak' else f"{sc['spp']} spp/step split over the GPUs"
    return {'workload': f"{sc['name']}: {len(sc['mtlids'])} tris, {nx}x{ny}, {per}, engine={sc['engine']}",
            'scene': sc['name'], 'tris': int(len(sc['mtlids'])), 'resolution': [nx, ny], 'spp_per_step_per_gpu': sc['spp'],
            'parallelism': f'sample-range x{n_gpus}' if n_gpus > 1 else 'single GPU',
            'l2_policy': 'path state and ray queues of a step (>= 0.9 GB) exceed the 126 MB L2, so every step starts cold'}


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

    def finish(self):
        self.stop_flag = True
        self.join(timeout=10)
        return self.summary()

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


# =====================================================================================================================
class Bench:
    """One process = one GPU.  Holds the distributed plumbing and the measurement primitives shared by every config."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world, self.local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        from ptina_b200 import scenes, worker, _native
        from ptina_b200 import dist as pdist
        self.scenes, self.worker, self.native, self.pdist = scenes, worker, _native, pdist
        worker.init(device=self.local)
        self.ctx = _native.context()
        self.current = None

    # ---- plumbing ----
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device='cuda', dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device='cuda', dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def load(self, sc):
        if self.current != sc['name']:
            self.scenes.apply(self.worker, sc)
            self.current = sc['name']
        self.ctx.sobol_reset()
        self.worker.clear()
        return {'path': self.native.ENGINE_PATH, 'brute': self.native.ENGINE_BRUTE, 'mlt': self.native.ENGINE_MLT}[sc['engine']]

    # ---- one path-traced step: `total` Sobol indices from the generator's time on, interleaved over the ranks, + film reduce ----
    def make_frame(self, eng, total, reduce=True):
        ctx, pdist, rank, world = self.ctx, self.pdist, self.rank, self.world

        def frame():
            k0 = ctx.sobol_time + 1
            first, count, stride = pdist.shard_range(k0, total, rank, world)
            if count:
                ctx.render_range(eng, first, count, stride)
            ctx.sobol_time = ctx.sobol_time + total
            if world > 1 and reduce:
                return pdist.reduce_film_pass(0, 0)
            return None
        return frame

    def count_rays(self, frame):
        ctx = self.ctx
        ctx.sobol_reset(); self.worker.clear()
        ctx.set_counting(True, False); ctx.reset_counters()
        frame(); ctx.synchronize()
        cnt = ctx.counters()
        ctx.set_counting(False, False)
        return cnt

    def timed(self, frame, steps, warmup, clocks=False):
        """device time of `steps` calls (max over ranks), host time spent enqueueing them, clocks during the window"""
        ctx = self.ctx
        ctx.sobol_reset(); self.worker.clear()
        for _ in range(warmup):
            frame()
        ctx.reset_counters()
        sampler = None
        if clocks:
            sampler = ClockSampler(self.local); sampler.start()
        ev0, ev1 = self.event(), self.event()
        self.barrier()
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            frame()
        ctx.flush()
        host_ms = (time.perf_counter() - t0) * 1e3
        ev1.record()
        self.barrier()
        ms = self.max_over_ranks(ev0.elapsed_time(ev1))
        out = {'ms': ms, 'host_enqueue_ms_per_step': host_ms / steps, 'launches': ctx.launches()}
        if sampler is not None:
            out['clocks'] = sampler.finish()
        return out

    def stage_pass(self, frame, nprof):
        ctx = self.ctx
        ctx.set_counting(False, True); ctx.reset_counters()
        for _ in range(nprof):
            frame()
        ctx.synchronize()
        stage = ctx.stage_ms()
        ctx.set_counting(False, False)
        return {k: v / nprof for k, v in stage.items()}

    def split_times(self, eng, total, reps=3):
        """render and reduce timed separately, max over ranks: what a sharded step is made of.  The ranks are re-aligned (device
        synchronise + barrier) between the two, so `reduce_ms` is the collective itself (staging copy + ncclReduce), not the wait for
        the slowest rank's render -- that wait is part of `render_ms` (max over ranks)."""
        ctx, pdist = self.ctx, self.pdist
        frame_nr = self.make_frame(eng, total, reduce=False)
        a, b, c, d = self.event(), self.event(), self.event(), self.event()
        r_ms = d_ms = 0.0
        for _ in range(reps):
            self.barrier()
            a.record(); frame_nr(); ctx.flush(); b.record()
            self.barrier()
            c.record()
            if self.world > 1:
                pdist.reduce_film_pass(0, 0)
            d.record()
            self.barrier()
            r_ms += self.max_over_ranks(a.elapsed_time(b)); d_ms += self.max_over_ranks(c.elapsed_time(d))
        return r_ms / reps, d_ms / reps

    def film_check(self, eng, total):
        """max relative difference between the reduced film of a sharded step and rank 0 rendering the same Sobol indices alone"""
        ctx, torch = self.ctx, self.torch
        ctx.sobol_reset(); self.worker.clear()
        k0 = ctx.sobol_time + 1
        red = self.make_frame(eng, total)()
        self.barrier()
        err = w_ok = None
        if self.rank == 0:
            red = red.clone()
            self.worker.clear()
            ctx.render_range(eng, k0, total, 1)
            one = ctx.film_tensor(0)
            scale = one[:, :3].abs().max().clamp_min(1e-6)
            err = float(((red[:, :3] - one[:, :3]).abs().max() / scale).item())
            w_ok = bool(torch.equal(red[:, 3], one[:, 3]) and float(one[:, 3].min().item()) == float(total))
        self.barrier()
        return {'max_rel_diff': err, 'sample_counts_equal': w_ok, 'samples': total, 'tolerance': 1e-5,
                'ok': (err is not None and err <= 1e-5 and w_ok) if self.rank == 0 else None}

    # ---- a path-traced config, end to end ----
    def measure_pt(self, sc, mode, steps, warmup, stages=False, clocks=False, check=True):
        eng = self.load(sc)
        spp, world = sc['spp'], self.world
        total = spp * world if mode == 'weak' else spp
        frame = self.make_frame(eng, total)
        cnt = self.count_rays(frame)
        rays_step = self.sum_over_ranks(cnt['rays'])
        t = self.timed(frame, steps, warmup, clocks=clocks)
        ms_step = t['ms'] / steps
        nx, ny = sc['size']
        out = {'mode': mode, 'n_gpus': world, 'resolution': [nx, ny], 'tris': int(len(sc['mtlids'])), 'engine': sc['engine'], 'spp_per_step_total': total,
               'Mrays_per_s': rays_step / (ms_step * 1e-3) / 1e6, 'spp_per_s': total / (ms_step * 1e-3), 'ms_per_step': ms_step, 'rays_per_step': rays_step,
               'host_enqueue_ms_per_step': t['host_enqueue_ms_per_step'], 'steps': steps, 'warmup': warmup,
               'per_ray_rank0': {'node_visits': cnt['node_visits'] / max(1, cnt['rays']), 'tri_tests': cnt['tri_tests'] / max(1, cnt['rays'])}}
        out['_raw'] = {'cnt': cnt, 'timed': t, 'frame': frame, 'eng': eng, 'total': total}
        if stages:
            out['stage_ms_per_step'] = self.stage_pass(frame, max(1, min(steps, 5)))
        if world > 1:
            r_ms, d_ms = self.split_times(eng, total)
            out['render_ms'], out['reduce_ms'] = r_ms, d_ms
            out['reduce_bytes'] = nx * ny * 16
            if check:
                out['film_check'] = self.film_check(eng, total)
        return out

    def limits(self, out, single_ms):
        """what keeps a sharded step from single_ms / N: smaller wavefronts (kernel tails, fewer rays per persistent kernel), the film
        reduce, or the host not enqueueing fast enough -- each in ms per step"""
        n = self.world
        ideal = single_ms / n
        tail = max(0.0, out.get('render_ms', out['ms_per_step']) - ideal)
        host_gap = max(0.0, out['host_enqueue_ms_per_step'] - out.get('render_ms', out['ms_per_step']))
        parts = {'ideal_ms (1-GPU time / N)': ideal, 'kernel_tail_ms (render - ideal)': tail, 'reduce_ms': out.get('reduce_ms', 0.0), 'host_enqueue_gap_ms': host_gap}
        worst = max(('kernel_tail_ms (render - ideal)', 'reduce_ms', 'host_enqueue_gap_ms'), key=lambda k: parts[k])
        return dict(parts, limiter=worst.split(' ')[0], single_gpu_ms=single_ms)

    def single_gpu_ms(self, sc, total, reps=2):
        """rank 0 alone renders the whole step (the others wait): the strong-scaling reference measured in the same run"""
        eng = self.load(sc)
        ms = 0.0
        if self.rank == 0:
            ctx = self.ctx
            ctx.render_range(eng, 65, total, 1); ctx.synchronize()
            a, b = self.event(), self.event()
            a.record()
            for _ in range(reps):
                ctx.render_range(eng, 65, total, 1)
            b.record(); self.torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
        self.barrier()
        return self.max_over_ranks(ms)

    # ---- config 5 ----
    def measure_mlt(self, calls=16, warmup=3):
        from ptina_b200.engine import MLTPathEngine
        sc = self.scenes.metropolis()
        self.load(dict(sc, engine='path'))
        nch = 1 << 18
        first, per = self.pdist.shard_chains(nch, self.rank, self.world)
        eng = MLTPathEngine(nchains=nch, seed=0, chains=(first, per))
        eng.chains = (first, per)
        eng.reset()
        ctx = self.ctx
        self.worker.clear()
        ctx.set_counting(True, False); ctx.reset_counters()
        eng.render(1); ctx.synchronize()
        rays_call = self.sum_over_ranks(ctx.counters()['rays'])
        ctx.set_counting(False, False)
        for _ in range(warmup):
            eng.render(1)
        a, b, c = self.event(), self.event(), self.event()
        self.barrier()
        t0 = time.perf_counter()
        a.record()
        for _ in range(calls):
            eng.render(1)
        host_ms = (time.perf_counter() - t0) * 1e3
        b.record()
        if self.world > 1:
            self.pdist.reduce_film_pass(0, 0)
        c.record()
        self.barrier()
        ms, red = self.max_over_ranks(a.elapsed_time(b)), self.max_over_ranks(b.elapsed_time(c))
        return {'mode': 'strong (2^18 chains split over the GPUs)', 'n_gpus': self.world, 'chains_total': nch, 'chains_per_gpu': per, 'dims': 32, 'render_calls': calls,
                'ms_per_render': ms / calls, 'Mproposals_per_s': nch * calls / (ms * 1e-3) / 1e6, 'Mrays_per_s': rays_call * calls / (ms * 1e-3) / 1e6,
                'host_enqueue_ms_per_render': host_ms / calls, 'reduce_ms_per_readback': red, 'resolution': list(sc['size']),
                'parity': 'chain trajectories unpinned (ti.random); per-chain radiance, accept/reject and splats are checked in tests/test_gpu_parity.py::test_mlt_engine'}


def issue_roofline(scene, spp_per_step, ms_per_step, sm_mhz):
    """What binds the small-scene configs is instruction issue (nodes in shared memory, triangles in L1).  Roof = 4 warp instructions
    per clock per SM x 148 SMs x the SM clock sampled during the timed region.  Warp instructions per step are a property of the
    deterministic workload: counted once per build with ncu (smsp__inst_executed.sum over every launch of one step: `bench.py --one-step`,
    tools/ncu_inst_counts.py -> profiles/inst_counts.json) and divided here by the LIVE step time."""
    sm_mhz = sm_mhz or 1965
    peak = 4.0 * 148 * sm_mhz * 1e6 / 1e9            # G warp-instructions / s
    out = {'bound': 'issue', 'unit': 'Gwarp-inst/s', 'peak': peak, 'peak_source': f'4 warp-instr/clk/SM x 148 SMs x {sm_mhz} MHz (SM clock sampled during the timed region)',
           'kernel': 'whole step (k_trace_pre + k_trace_tree + k_shade are 97 % of it)', 'timing': 'CUDA events over the timed region (ms_per_step)'}
    inst = {}
    try:
        inst = json.load(open(os.path.join(ROOT, 'profiles', 'inst_counts.json'))).get(scene, {})
    except Exception:
        pass
    if inst.get('warp_inst_per_step') and inst.get('spp'):
        w = inst['warp_inst_per_step'] * spp_per_step / inst['spp']         # instructions scale with the samples of a step
        ach = w / (ms_per_step * 1e-3) / 1e9
        lanes = inst['thread_inst_per_step'] / inst['warp_inst_per_step']
        out.update({'achieved': ach, 'frac': ach / peak, 'warp_inst_per_step': w, 'lanes_per_inst': lanes, 'frac_lane_weighted': ach / peak * lanes / 32.0,
                    'inst_source': inst.get('source'), 'traffic': inst.get('dram_bytes_per_step', 0) * spp_per_step / inst['spp'], 'per_kernel': inst.get('per_kernel')})
    else:
        out.update({'achieved': None, 'frac': None, 'traffic': None, 'note': 'profiles/inst_counts.json has no entry for this scene'})
    return out


def run_gpu(args):
    B = Bench(args)
    torch, ctx, worker, native = B.torch, B.ctx, B.worker, B.native
    rank, world = B.rank, B.world
    sc = workload(args)
    nx, ny = sc['size']
    spp = sc['spp']

    if args.one_step:           # ncu --profile-from-start off: exactly one warm step between cudaProfilerStart / Stop
        eng = B.load(sc)
        frame = B.make_frame(eng, spp * world)
        for _ in range(max(1, args.warmup)):
            frame()
        ctx.synchronize()
        torch.cuda.profiler.start()
        frame(); ctx.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({'one_step': sc['name'], 'spp': spp}))
        return

    head = B.measure_pt(sc, 'weak', args.steps, args.warmup, stages=True, clocks=True, check=not args.quick)
    raw = head.pop('_raw')
    cnt, frame, eng = raw['cnt'], raw['frame'], raw['eng']
    ms = raw['timed']['ms']
    launches = raw['timed']['launches']
    clocks = raw['timed']['clocks']
    stage = head['stage_ms_per_step']
    total_rays_per_step = head['rays_per_step']
    value = head['Mrays_per_s']
    info = ctx.tree

    # ---- sustained window: the same step for >= 2 s ----
    sustained = None
    if not args.quick:
        n_sus = max(args.steps, int(2200.0 / max(head['ms_per_step'], 1e-3)) + 1)
        t = B.timed(frame, n_sus, 1, clocks=True)
        sustained = {'value': total_rays_per_step * n_sus / (t['ms'] * 1e-3) / 1e6, 'unit': UNIT, 'seconds': t['ms'] * 1e-3, 'steps': n_sus,
                     'ms_per_step': t['ms'] / n_sus, 'clocks': t['clocks']}

    # ---- end to end through the public API with HOST buffers (pinned), H2D + build + render + D2H inside the timed region ----
    verts_pin = torch.from_numpy(np.ascontiguousarray(sc['vertices'], dtype=np.float32)).pin_memory()
    mtl_pin = torch.from_numpy(np.ascontiguousarray(sc['mtlids'], dtype=np.int32)).pin_memory()
    img_pin = torch.empty((nx, ny, 4), dtype=torch.float32).pin_memory()
    from ptina_b200.model import ModelPool
    from ptina_b200.tree import BVHTree
    from ptina_b200.engine import PathEngine, BruteEngine
    engine_cls = BruteEngine if sc['engine'] == 'brute' else PathEngine

    def e2e_run(percall):
        parts = {}

        def step():
            t = [time.perf_counter()]
            ModelPool().load(verts_pin.numpy(), mtl_pin.numpy()); t.append(time.perf_counter())
            BVHTree().build(); t.append(time.perf_counter())
            worker.clear()
            tot = None
            if percall:
                for _ in range(spp):              # exams/benchmark.py:29-33: one sample per call
                    engine_cls().render()
            else:
                tot = frame()
            t.append(time.perf_counter())
            if world > 1:                          # rank 0 resolves the reduced film (rgb / w) and reads it back; the others drain
                if rank == 0:
                    img_pin.copy_(B.pdist.resolve(tot))
                else:
                    ctx.synchronize()
            else:
                ctx.get_image(0, out=img_pin.numpy())
            t.append(time.perf_counter())
            for name, a, b in zip(('load', 'build', 'launch', 'wait+readback'), t[:-1], t[1:]):
                parts.setdefault(name, []).append(round((b - a) * 1e3, 3))
        return step, parts

    def e2e_measure(percall):
        step, parts = e2e_run(percall)
        for _ in range(max(1, args.warmup)):
            step()
        B.barrier()
        t0 = time.perf_counter()
        e0, e1 = B.event(), B.event()
        e0.record()
        per = []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            step()
            per.append((time.perf_counter() - t1) * 1e3)
        e1.record()
        B.barrier()
        tot_ms = B.max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
        return {'value': total_rays_per_step * args.steps / (tot_ms * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': tot_ms / args.steps,
                'ms_steps': [round(x, 3) for x in per], 'parts_ms_last_step': {k: v[-1] for k, v in parts.items()}}

    if world > 1:
        # percall at N > 1 is not a reference pattern (the reference is single-device): batch only
        e2e = e2e_measure(False)
        e2e_percall = None
    else:
        e2e = e2e_measure(False)
        e2e_percall = e2e_measure(True)
    e2e.update({'h2d_bytes_per_step': int(verts_pin.numel() * 4 + mtl_pin.numel() * 4), 'd2h_bytes_per_step': int(img_pin.numel() * 4),
                'path': 'ModelPool.load(pinned host) + BVHTree.build + FilmTable.clear + render_range(32 spp' + (', sharded) + NCCL film reduce + resolve' if world > 1 else ')') + ' + get_image(host)'})
    if e2e_percall:
        e2e_percall.update({'h2d_bytes_per_step': e2e['h2d_bytes_per_step'], 'd2h_bytes_per_step': e2e['d2h_bytes_per_step'], 'calls_per_step': spp,
                            'ratio_to_batch_e2e': e2e_percall['value'] / e2e['value'],
                            'path': f'ModelPool.load + BVHTree.build + FilmTable.clear + {spp} x {engine_cls.__name__}().render() + get_image(host)  (exams/benchmark.py:29-34)'})

    # ---- the other configs at this N ----
    configs = None
    if not args.quick:
        configs = {}
        steps_small = max(3, min(args.steps, 5))
        head_cfg = {k: v for k, v in head.items()}
        head_cfg['e2e_Mrays_per_s'] = e2e['value']
        configs['config2_cornell_monkey'] = head_cfg
        for key, name in (('config1_cornell_boxes', 'cornell_boxes'), ('config3_matball', 'matball')):
            if name == sc['name']:
                continue
            r = B.measure_pt(workload(args, name), 'weak', steps_small, 2, stages=True)
            r.pop('_raw')
            configs[key] = r
        # config 2 in PTB_MODE_FAST (opt-in, NON-parity: approximate division / FMA contraction in the shading stage only)
        ctx.set_mode('fast')
        r = B.measure_pt(workload(args, 'cornell_monkey'), 'weak', steps_small, 2, stages=True, check=False)
        r.pop('_raw')
        r['note'] = 'non-parity mode (ptb_set_mode PTB_MODE_FAST): holds the 1e-3 image gate, not the 1e-5 BSDF tap gate; not the headline'
        configs['config2_fast_mode_nonparity'] = r
        ctx.set_mode('parity')
        # config 2, strong: ONE 32-spp frame split over the N GPUs
        if world > 1:
            s2 = workload(args, 'cornell_monkey')
            single = B.single_gpu_ms(s2, s2['spp'])
            r = B.measure_pt(s2, 'strong', steps_small, 2)
            r.pop('_raw')
            r['limits'] = B.limits(r, single)
            configs['config2_strong_one_frame_split'] = r
        # config 4: a 32-spp step of the 1080p frame split over the N GPUs + the 33 MB reduce; then the whole 1024-spp render
        s4 = workload(args, 'mega')
        single4 = B.single_gpu_ms(s4, s4['spp'], reps=1) if world > 1 else None
        r = B.measure_pt(s4, 'strong', 3, 1, stages=(world == 1))
        raw4 = r.pop('_raw')
        if world > 1:
            r['limits'] = B.limits(r, single4)
        c4, st4 = raw4['cnt'], r.get('stage_ms_per_step')
        if st4:
            # L2 roofline of the traversal stages, SURVEY 8(d): B_ray = 64 N_int + 48 N_tri + 48 bytes
            trav_ms = st4['extend'] + st4['shadow']
            alg = 64.0 * c4['node_visits'] + 48.0 * c4['tri_tests'] + 48.0 * c4['rays']
            l2 = ctx.measure_l2(64, 20)
            wide4 = os.environ.get('PTB_NO_WIDE4') is None       # trees walked out of global memory use 4-wide 64-byte nodes unless switched off
            fmt = (64.0 if wide4 else 32.0) * c4['node_visits'] + 64.0 * c4['tri_tests'] + 48.0 * c4['rays']
            r['roofline_l2'] = {'bound': 'l2', 'achieved': alg / (trav_ms * 1e-3) / 1e9, 'peak': l2, 'unit': 'GB/s', 'frac': alg / (trav_ms * 1e-3) / 1e9 / l2,
                                'formula': 'SURVEY 8(d): 64 B x node visits + 48 B x triangle tests + 48 B x rays, counters_per_step / (extend + shadow stage time, serial pass)',
                                'bytes_in_this_format': fmt, 'format_note': ('a node visit is one 4-wide quantised node: 64 B, two 256-bit loads, up to four child boxes' if wide4 else 'nodes are fetched as 32-B quantised Node32 (one 256-bit load)') + '; triangles as 64-B Tri64',
                                'note': 'not the binding roof: the kernel waits on dependent fetches and on instruction issue (roofline_issue next to this); wider nodes and a better tree LOWER this numerator (fewer visits per ray) while the frame gets faster',
                                'counters_per_step': {k: c4[k] for k in ('rays', 'node_visits', 'tri_tests')}, 'traversal_ms_per_step': trav_ms,
                                'peak_source': 'ptb_measure_l2: 148 x 8 blocks x 256 threads stream a 64 MiB buffer 20 times with 16-byte ld.global.cg loads after 2 warm passes (CUDA events)'}
        # the full config: 1024 spp, Sobol points 65..1088, sharded, one reduce at the end
        eng4 = B.load(s4)
        full_frame = B.make_frame(eng4, 1024)
        samp = ClockSampler(B.local); samp.start()
        a, b = B.event(), B.event()
        B.barrier(); a.record(); full_frame(); ctx.flush(); b.record(); B.barrier()
        full_ms = B.max_over_ranks(a.elapsed_time(b))
        r['full_1024spp'] = {'seconds': full_ms * 1e-3, 'spp_per_s': 1024 / (full_ms * 1e-3), 'Mrays_per_s': r['rays_per_step'] / r['spp_per_step_total'] * 1024 / (full_ms * 1e-3) / 1e6,
                             'includes': 'render of 1024 spp split over the GPUs + one film reduce', 'clocks': samp.finish()}
        configs['config4_mega'] = r
        configs['config5_metropolis'] = B.measure_mlt()
        B.load(sc)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks['hbm_gbs'], 'measured (MEASURED_PEAKS.json)') if 'hbm_gbs' in peaks else (6650.0, 'fallback (B200_PROFILING.md)')
        roofline = issue_roofline(sc['name'], spp, head['ms_per_step'], clocks.get('sm_mhz'))
        if configs is not None:
            for key, name in (('config1_cornell_boxes', 'cornell_boxes'), ('config3_matball', 'matball')):
                if key in configs:
                    r = issue_roofline(name, configs[key]['spp_per_step_total'] / world, configs[key]['ms_per_step'], clocks.get('sm_mhz'))   # per GPU
                    r.pop('per_kernel', None)
                    configs[key]['roofline'] = r
            if 'config4_mega' in configs:
                # config 4 next to its L2 roofline: the kernel is bound by dependent-fetch latency and instruction issue, not by L2 bytes
                # (a 4-wide step fetches 64 B for what two or three binary steps fetched before, so the byte numerator FALLS as the kernel
                # gets faster); the issue-slot view uses the instruction count of a 4-spp step (profiles/inst_counts.json), scaled by spp
                r = issue_roofline('mega', configs['config4_mega']['spp_per_step_total'] / world, configs['config4_mega']['ms_per_step'], clocks.get('sm_mhz'))
                r.pop('per_kernel', None)
                configs['config4_mega']['roofline_issue'] = r
        # the schema's HBM view of the traversal stages, with SURVEY 8(d)'s formula (served from shared memory / L1, so not the binding roof)
        trav_ms = stage['extend'] + stage['shadow']
        alg = 64.0 * cnt['node_visits'] + 48.0 * cnt['tri_tests'] + 48.0 * cnt['rays']
        roofline['traversal_bytes_view'] = {'algorithmic_GBps': alg / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else None, 'hbm_peak_GBps': hbm_peak, 'peak_source': peak_src,
                                            'formula': '64 B x node visits + 48 B x triangle tests + 48 B x rays (SURVEY 8d) / (extend + shadow stage time)',
                                            'note': 'not a roofline for this config: the 125 KB scene is served from shared memory and L1; DRAM traffic is `traffic`'}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': config_of(sc, world),
                'spp_per_s': spp * world * args.steps / (ms * 1e-3),
                'rays_per_step': total_rays_per_step,
                'e2e': e2e, 'e2e_percall': e2e_percall, 'sustained': sustained,
                'gpu_launches': int(launches),
                'roofline': roofline,
                'stage_ms_per_step': stage,
                'schedule': 'shadow stage of bounce b overlapped with the extend stage of bounce b+1 (2 streams); stage_ms_per_step from a serial pass',
                'counters_per_step_rank0': cnt,
                'tree': {'n': info.n, 'depth': info.depth, 'valid': info.valid, 'policy': info.policy, 'build_ms': info.build_ms, 'trav_depth': info.trav_depth, 'trav_ploc': info.trav_ploc},
                'clocks': clocks}
        for k in ('film_check', 'render_ms', 'reduce_ms', 'host_enqueue_ms_per_step'):
            if k in head:
                line[k] = head[k]
        if configs is not None:
            line['configs'] = configs
        if not args.no_cpu and world == 1:      # rank 0 at N=1 only (torchrun also pins OMP_NUM_THREADS=1)
            line['cpu_baseline'] = {k: v for k, v in cpu_reference(sc, args.cpu_seconds, max_frames=8).items() if k in ('value', 'unit', 'cores', 'kind', 'sample', 'spp_per_s', 'note')}
        print(json.dumps(line), flush=True)
    if world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)
