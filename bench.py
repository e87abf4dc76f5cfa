#!/usr/bin/env python3
"""bench.py -- Mpaths/s (and Mrays/s) of the hair path-tracing hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scene NAME] [--impl cuda|reference]

A step = one pass of the hot path over one batch of synthetic input = one full render of the workload
(default: BASELINE.json configs[1]: hair-curl, Marschner R/TT/TRT, 1024x1024 at 64 spp, maxDepth 65; procedural fibers
because the reference's .mitshair blobs are missing).  One JSON line is printed by rank 0.

  value      whole-job Mpaths/s with the flattened scene, BVH and tables already resident in HBM; device time (CUDA events on
             the render stream), max over ranks.  At N GPUs every rank renders its own `spp` sample indices of an N*spp-sample
             image (weak scaling) and the per-GPU films are summed with one NCCL reduce inside the timed region.
  e2e        same metric through the C ABI with HOST buffers: context creation, geometry/envmap upload from pinned host memory,
             BVH build, render, film read-back -- all inside the timed region.
  roofline   k_intersect (closest-hit BVH traversal, the dominant kernel): algorithmic bytes per launch over its measured
             launch time, against the measured HBM copy bandwidth.
  cpu_baseline  the CPU oracle (restatement of the reference, built with the reference's flags) on a bounded sample.

--impl reference times the CPU oracle with all host threads on the same config (bounded sample per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

import numpy as np


def load_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.samples = []; self.stop_flag = False; self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append((time.time(), [x.strip() for x in line.split(',')]))
        except Exception:
            pass

    def finish(self, t_begin=None, t_end=None):
        """Median SM clock / throttle reasons of the samples taken inside [t_begin, t_end] (the timed region); when the region is
        shorter than the sampling period, of the samples taken under load since the sampler started (warm-up included)."""
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        inside = [s for (t, s) in self.samples if t_begin is not None and t_begin <= t <= t_end + 0.25]
        for s in (inside if len(inside) >= 2 else [s for (_, s) in self.samples]):
            try:
                sm.append(float(s[1])); mx.append(float(s[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        load = sorted(sm)[len(sm) // 2:]   # upper half = samples under load
        return {'sm_mhz': float(np.median(load)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons)}


def scene_arrays(name, scale):
    """Flattened host inputs of the workload: per shape (xyz, starts, radius, bsdf props) + baked envmap (pinned where torch is available)."""
    import cudapath
    sc = cudapath.scenes.SCENES[name]
    shapes = []
    for sh in sc['shapes']:
        xyz, st = cudapath.scenes.generate(sh, scale)
        shapes.append((xyz, st, sh['radius'], dict(sh['bsdf'])))
    env = cudapath.bake_sunsky(**cudapath.scenes.sunsky_params(name))
    return sc, shapes, env


def pin(arr):
    """Copies a numpy array into page-locked host memory (the e2e path uploads from pinned memory)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t.numpy(), t


def make_context(cudapath, sc, shapes, env, device):
    ctx = cudapath.Context(device)
    for xyz, st, radius, b in shapes:
        b = dict(b); t = b.pop('type'); b.pop('id', None)
        ctx.add_hair(xyz, st, radius, ctx.add_bsdf(t, **b))
    ctx.set_envmap(env)
    ctx.set_camera(np.array(sc['camera'], np.float32).reshape(4, 4), sc['fov'], width=sc['width'], height=sc['height'])
    ctx.set_film('tent')
    ctx.set_integrator(maxDepth=sc['maxDepth'], rrDepth=5, strictNormals=True)
    return ctx


def run_reference(args, rank, world):
    """CPU arm: the oracle (kind "port": the reference cannot be built here) with all host threads."""
    if rank != 0:
        return
    import cudapath
    import orc
    import ctypes
    fast = os.path.join(REPO, 'oracle', 'liboracle_fast.so')
    if os.path.exists(fast):            # reference's optimisation flags
        orc._lib = None; orc.ORACLE_LIB = fast
    sc, shapes, env = scene_arrays(args.scene, args.scale)
    t0 = time.time()
    osc = orc.Scene()
    for xyz, st, radius, b in shapes:
        b = dict(b); t = b.pop('type'); b.pop('id', None)
        osc.add_hair(xyz, st, radius, osc.add_bsdf(t, **b))
    osc.set_envmap(env)
    osc.set_camera(np.array(sc['camera'], np.float32).reshape(4, 4), sc['fov'], width=sc['width'], height=sc['height'])
    osc.set_film('tent'); osc.set_integrator(maxDepth=sc['maxDepth'], rrDepth=5, strictNormals=True)
    osc.build()
    build_s = time.time() - t0
    cores = os.cpu_count() or 1
    spp = sc['spp']
    # bounded sample: the first `cpu_spp` sample indices of every pixel (paths/s is independent of the sample index)
    cpu_spp = max(1, min(spp, args.cpu_spp))
    for _ in range(args.warmup if args.impl == 'reference' else 0):
        osc.render(spp, seed=1, sample_begin=0, sample_end=1, threads=cores)
    times, rays = [], 0
    steps = args.steps if args.impl == 'reference' else 1
    for k in range(steps):
        t = time.time()
        osc.render(spp, seed=1 + k, sample_begin=0, sample_end=cpu_spp, threads=cores)
        times.append(time.time() - t)
        rays = osc.last_stats['rays'] + osc.last_stats['shadow_rays']
    paths = sc['width'] * sc['height'] * cpu_spp
    mean_t = float(np.mean(times))
    val = paths / mean_t / 1e6
    base = {'value': val, 'unit': 'Mpaths/s', 'cores': cores, 'kind': 'port',
            'sample': '%s: all %dx%d pixels x first %d of %d samples (%d paths, %.1f s per step), oracle built with the reference flags; kd-tree build replaced by a binned-SAH BVH (%.1f s, untimed)'
                      % (args.scene, sc['width'], sc['height'], cpu_spp, spp, paths, mean_t, build_s),
            'mrays_per_s': rays / mean_t / 1e6}
    if args.impl != 'reference':
        return base
    line = {'impl': 'reference', 'metric': 'Mpaths/s', 'value': val, 'unit': 'Mpaths/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
            'ms_per_step': mean_t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, sc), 'cpu_baseline': base,
            'e2e': {'value': val, 'unit': 'Mpaths/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'mrays_per_s': base['mrays_per_s']}
    print(json.dumps(line), flush=True)


def workload_config(args, sc):
    return {'workload': '%s: %s BSDF, path integrator maxDepth=%d rrDepth=5 strictNormals, %dx%d at %d spp per GPU, sunsky envmap 512x256, tent filter'
                        % (args.scene, sc['shapes'][0]['bsdf']['type'], sc['maxDepth'], sc['width'], sc['height'], sc['spp']),
            'geometry': 'procedural fibers (reference .mitshair blobs missing), generator scale %.3g' % args.scale,
            'parallelism': 'sample-range sharding, one NCCL film reduce', 'l2_policy': 'inputs larger than L2 (BVH + vertices + path queues >> 126 MB)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--scene', default='hair-curl', choices=['straight-hair', 'hair-curl', 'curly-hair', 'furball'])
    ap.add_argument('--scale', type=float, default=1.0, help='strand-count scale of the procedural generators')
    ap.add_argument('--spp', type=int, default=0, help='override samples per pixel (default: the config value)')
    ap.add_argument('--cpu-spp', type=int, default=1, help='sample indices per pixel in the bounded CPU sample')
    ap.add_argument('--wave', type=int, default=0)
    ap.add_argument('--max-split', type=int, default=0)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    import cudapath
    if args.spp:
        cudapath.scenes.SCENES[args.scene] = dict(cudapath.scenes.SCENES[args.scene], spp=args.spp)
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    sc, shapes, env = scene_arrays(args.scene, args.scale)
    W, H, spp = sc['width'], sc['height'], sc['spp']
    total_spp = spp * world                         # weak scaling: every rank renders `spp` sample indices of a world*spp image
    ctx = make_context(cudapath, sc, shapes, env, local)
    if args.max_split:
        ctx.set_build_options(args.max_split)
    if args.wave:
        ctx.set_options(wave_size=args.wave)
    ctx.build()
    build = ctx.stats()
    # a real (non-default) stream shared by torch (film zeroing, NCCL reduce, timing events) and the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    film = torch.zeros((H, W, 5), dtype=torch.float32, device='cuda')

    def step(seed):
        film.zero_()
        ctx.render_into(film.data_ptr(), total_spp, seed=seed, sample_begin=rank * spp, sample_end=(rank + 1) * spp, stream=stream.cuda_stream)
        if world > 1:
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)

    sampler = ClockSampler(local); sampler.start()
    for w in range(args.warmup):
        step(100 + w)
    # one profiled pass (stage timing by CUDA events around every launch) and one counting pass (nodes / primitives) -- both untimed
    ctx.set_options(wave_size=args.wave, profile_stages=True)
    step(7); torch.cuda.synchronize(); prof = ctx.stats()
    ctx.set_options(wave_size=args.wave, collect_stats=True)
    step(7); torch.cuda.synchronize(); cnt = ctx.stats()
    ctx.set_options(wave_size=args.wave)

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = time.time()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    launches = 0; rays = 0; shadow = 0
    e0.record(stream)
    for k in range(args.steps):
        step(1000 + k)
        st = ctx.stats(); launches += st['kernel_launches'] + 1; rays += st['rays']; shadow += st['shadow_rays_traced']      # rays actually traced (zero-contribution shadow rays are only counted)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.finish(t_begin, time.time())
    t = torch.tensor([ms, float(rays), float(shadow), float(launches)], dtype=torch.float64, device='cuda')
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0]); rays = float(tsum[1]); shadow = float(tsum[2]); launches = int(tsum[3])
    ms_per_step = ms / args.steps
    paths_per_step = W * H * spp * world
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6
    mrays = (rays + shadow) / args.steps / (ms_per_step * 1e-3) / 1e6

    # ---- e2e through the C ABI with host buffers (pinned), everything inside the timed region
    e2e = None
    if not args.no_e2e:
        pinned = []; pshapes = []
        for xyz, stf, radius, b in shapes:
            a, ta = pin(xyz); s2, ts = pin(stf); pinned += [ta, ts]; pshapes.append((a, s2, radius, b))
        penv, te = pin(env); pinned.append(te)
        h2d = sum(16 * (len(s[1]) + 1) for s in pshapes) + penv.nbytes      # float4 vertex stream (+ sentinel) + envmap fp32
        d2h = W * H * 5 * 4
        # W untimed warm-up runs (the first ones grow the stream-ordered memory pool and find the GPU at idle clocks), then K timed runs
        times = []; phases = []
        n_warm = max(args.warmup, 0); n_timed = max(args.steps, 1)
        for k in range(n_warm + n_timed):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            c2 = make_context(cudapath, sc, pshapes, penv, local)
            if args.wave:
                c2.set_options(wave_size=args.wave)
            t1 = time.perf_counter()
            c2.build()
            t2 = time.perf_counter()
            c2.render(total_spp, seed=2000 + k, sample_begin=rank * spp, sample_end=(rank + 1) * spp)
            torch.cuda.synchronize()
            if k >= n_warm:
                times.append(time.perf_counter() - t0)
                phases.append([round(t1 - t0, 4), round(t2 - t1, 4), round(time.perf_counter() - t2, 4), round(c2.stats()['render_ms'] * 1e-3, 4)])
            c2.close()
        te2e = torch.tensor([float(np.mean(times))], dtype=torch.float64, device='cuda')       # mean of the K timed runs
        if world > 1:
            dist.all_reduce(te2e, op=dist.ReduceOp.MAX)
        e2e = {'times_s': [round(t, 4) for t in times], 'phases_s(create+upload, build, render+readback, of which device render)': phases, 'value': paths_per_step / float(te2e[0]) / 1e6, 'unit': 'Mpaths/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'warmup_runs': n_warm, 'includes': 'context creation, Marschner table build, geometry+envmap upload, device BVH build, render, film read-back'}

    if rank == 0:
        peak, peak_kind = load_peaks()
        # k_intersect: algorithmic bytes = 40 B/ray (32 in + 8 out) + 128 B per node visited + 52 B per primitive tested
        n_launch = max(1, prof['intersect_launches'])
        alg_bytes = 40.0 * cnt['rays'] + 128.0 * cnt['nodes_visited'] + 52.0 * cnt['prims_tested']
        achieved = alg_bytes / (prof['intersect_ms'] * 1e-3) / 1e9 if prof['intersect_ms'] > 0 else 0.0
        # DRAM bytes of ONE captured k_intersect launch (ncu --set full, profiles/k_intersect_traffic.json written by tools/summarize_profile.py);
        # bench.py itself never runs under a profiler
        traffic = traffic_note = None
        tp = os.path.join(REPO, 'profiles', 'k_intersect_traffic.json')
        if os.path.exists(tp) and args.scene == 'hair-curl':
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj['dram_bytes']; traffic_note = '%s; that launch ran %.3f ms under ncu' % (tj['launch'], tj['duration_ms'])
        roof = {'kernel': 'k_intersect (closest-hit BVH4 traversal + FP64 cylinder test)', 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'traffic_note': traffic_note, 'peak_source': peak_kind,
                'bytes_per_launch': alg_bytes / n_launch, 'avg_launch_ms': prof['intersect_ms'] / n_launch, 'launches_per_step': n_launch,
                'nodes_per_ray': cnt['nodes_visited'] / max(1, cnt['rays']), 'prims_per_ray': cnt['prims_tested'] / max(1, cnt['rays']),
                'exact_tests_per_ray': cnt['full_tests'] / max(1, cnt['rays']),
                'stage_share_of_step': {k: prof[k + '_ms'] / max(prof['render_ms'], 1e-9) for k in ('intersect', 'shade', 'shadow', 'raygen', 'splat')}}
        cpu = None
        if not args.no_cpu and world == 1:
            cpu = run_reference(args, 0, 1)
        line = {'metric': 'Mpaths/s', 'value': value, 'unit': 'Mpaths/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (fp64 cylinder test)', 'data': 'synthetic',
                'config': workload_config(args, sc), 'mrays_per_s': mrays, 'rays_per_path': (rays + shadow) / args.steps / paths_per_step,
                'gpu_launches': int(launches), 'clocks': clocks, 'e2e': e2e, 'roofline': roof, 'cpu_baseline': cpu,
                'build': {'segments': build['segments'], 'bvh_references': build['bvh_references'], 'bvh_nodes': build['bvh_nodes'], 'build_ms': build['build_ms'], 'build_ms_note': 'first build of the process (cold device allocator); the steady-state build is the second entry of e2e.phases_s'}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
