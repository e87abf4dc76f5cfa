#!/usr/bin/env python3
"""bench.py -- Mpaths/s (and Mrays/s) of the hair path-tracing hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scene NAME] [--impl cuda|reference] [--scaling strong|weak] [--config 5]

A step = one pass of the hot path over one batch of synthetic input = one full render of the workload
(default: BASELINE.json configs[1]: hair-curl, Marschner R/TT/TRT, 1024x1024 at 64 spp, maxDepth 65; procedural fibers
because the reference's .mitshair blobs are missing).  One JSON line is printed by rank 0.

  value      whole-job Mpaths/s with the flattened scene, BVH and tables already resident in HBM; device time (CUDA events on
             the render stream), max over ranks.  At N GPUs the SAME job is split (default `--scaling strong`: the sample indices of
             the fixed 1024x1024 x 64 spp image are sharded over the ranks, north_star's "sample ranges split across 1/2/4/8 GPUs")
             and the per-GPU films are summed with one NCCL reduce inside the timed region.  `--scaling weak` renders 64 spp PER GPU.
  e2e        same metric through the C ABI with HOST buffers: context creation, geometry/envmap upload from pinned host memory,
             BVH build, render, the NCCL film reduce (N > 1) and the film read-back -- all inside the timed region.
  roofline   the dominant kernel (k_trace: BVH traversal of closest-hit + occlusion rays) against the resource that binds it, and one
             entry per stage under roofline.stages (see DESIGN.md section 3 for the byte / flop models).
  cpu_baseline  the CPU oracle (restatement of the reference, built with the reference's flags) on a bounded sample.

--impl reference times the CPU oracle with all host threads on the same config (bounded sample per step).
--config 5 prints the per-stage micro-benchmark line of BASELINE.json configs[4] (2^26 Marschner tuples, 2^26 rays vs the furball BVH).
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

HAIR_RGB = (0.143016, 0.0156076, 1.80928e-005)


def load_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j['hbm_gbs']), float(j.get('sm_max_mhz', 1965.0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1965.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    period_ms = int(os.environ.get('BENCH_CLOCK_PERIOD_MS', '200'))     # the profiling recipe's 200 ms; 0 = no sampling (development)

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.samples = []; self.stop_flag = False; self.proc = None

    def run(self):
        if self.period_ms <= 0:
            return
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', str(self.period_ms)],
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
        pw = []
        for s in (inside if len(inside) >= 2 else [s for (_, s) in self.samples]):
            try:
                pw.append(float(s[3]))
            except Exception:
                pass
        return {'sm_mhz': float(np.median(load)), 'sm_max_mhz': float(max(mx)), 'sm_min_mhz': float(min(sm)), 'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def scene_arrays(name, scale):
    """Flattened host inputs of the workload: per shape (xyz, starts, radius, bsdf props) + baked envmap (pinned where torch is available)."""
    import cudapath
    sc = cudapath.scenes.SCENES[name]
    cache = os.environ.get('CUDAPATH_SCENE_CACHE')          # development only: repeated A/B runs on one box skip the procedural generators
    cfile = os.path.join(cache, '%s_%g.npz' % (name, scale)) if cache else None
    if cfile and os.path.exists(cfile):
        z = np.load(cfile)
        shapes = [(z['xyz%d' % i], z['st%d' % i], sh['radius'], dict(sh['bsdf'])) for i, sh in enumerate(sc['shapes'])]
        return sc, shapes, z['env']
    shapes = []
    for sh in sc['shapes']:
        xyz, st = cudapath.scenes.generate(sh, scale)
        shapes.append((xyz, st, sh['radius'], dict(sh['bsdf'])))
    env = cudapath.bake_sunsky(**cudapath.scenes.sunsky_params(name))
    if cfile:
        os.makedirs(cache, exist_ok=True)
        np.savez(cfile, env=env, **{'xyz%d' % i: s[0] for i, s in enumerate(shapes)}, **{'st%d' % i: s[1] for i, s in enumerate(shapes)})
    return sc, shapes, env


def pin(arr):
    """Copies a numpy array into page-locked host memory (the e2e path uploads from pinned memory)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t.numpy(), t


def make_context(cudapath, sc, shapes, env, device, paths_per_device=None):
    ctx = cudapath.Context(device)
    if paths_per_device:
        ctx.set_job_size_hint(paths_per_device)     # build effort from the size of the job (cudapath.h)
    for xyz, st, radius, b in shapes:
        b = dict(b); t = b.pop('type'); b.pop('id', None)
        ctx.add_hair(xyz, st, radius, ctx.add_bsdf(t, **b))
    ctx.set_envmap(env)
    ctx.set_camera(np.array(sc['camera'], np.float32).reshape(4, 4), sc['fov'], width=sc['width'], height=sc['height'])
    ctx.set_film('tent')
    ctx.set_integrator(maxDepth=sc['maxDepth'], rrDepth=5, strictNormals=True)
    return ctx


def sample_range(total_spp, rank, world):
    import cudapath
    return cudapath.dist.sample_range(total_spp, rank, world)        # the host-side split the gloo test covers (tests/test_dist_gloo.py)


def run_reference(args, rank, world):
    """CPU arm: the oracle (kind "port": the reference cannot be built here) with all host threads.  The CPU build uses
    the reference's optimisation flags and the platform libm; what accelerates its ray queries is reported in `sample`."""
    if rank != 0:
        return
    import cudapath
    import orc
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
    accel = orc.accel_description() if hasattr(orc, 'accel_description') else 'binned-SAH BVH'
    base = {'value': val, 'unit': 'Mpaths/s', 'cores': cores, 'kind': 'port',
            'sample': '%s: all %dx%d pixels x first %d of %d samples (%d paths, %.1f s per step); oracle built with the reference flags; ray queries: %s, built in %.1f s (untimed)'
                      % (args.scene, sc['width'], sc['height'], cpu_spp, spp, paths, mean_t, accel, build_s),
            'mrays_per_s': rays / mean_t / 1e6}
    if args.impl != 'reference':
        return base
    line = {'impl': 'reference', 'metric': 'Mpaths/s', 'value': val, 'unit': 'Mpaths/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
            'ms_per_step': mean_t * 1e3, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, sc), 'cpu_baseline': base,
            'e2e': {'value': val, 'unit': 'Mpaths/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'mrays_per_s': base['mrays_per_s']}
    print(json.dumps(line), flush=True)


def workload_config(args, sc):
    per = 'per GPU' if args.scaling == 'weak' else 'in total, sample indices split over the GPUs'
    return {'workload': '%s: %s BSDF, path integrator maxDepth=%d rrDepth=5 strictNormals, %dx%d at %d spp %s, sunsky envmap 512x256, tent filter'
                        % (args.scene, sc['shapes'][0]['bsdf']['type'], sc['maxDepth'], sc['width'], sc['height'], sc['spp'], per),
            'geometry': 'procedural fibers (reference .mitshair blobs missing), generator scale %.3g' % args.scale,
            'parallelism': '%s sharding (%s scaling), one NCCL film reduce' % ('sample-range' if args.scaling == 'weak' or args.shard == 'samples' else '32x32 pixel-block', args.scaling),
            'l2_policy': 'inputs larger than L2 (BVH + vertices + path queues >> 126 MB)'}


def ncu_evidence():
    """Static evidence of the last committed ncu --set full capture of k_trace (profiles/k_trace_ncu.json, written by
    tools/summarize_profile.py); bench.py itself never runs under a profiler."""
    p = os.path.join(REPO, 'profiles', 'k_trace_ncu.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def stage_rooflines(prof, cnt, hbm_peak, l2_peak, sm_mhz, peak_src, scene):
    """One roofline entry per stage of the step.  Times: the profiled pass (CUDA events around every launch on the render stream).
    Counts: the counting pass of the same workload.  Models:
      k_trace   bytes the kernel LOADS through L1/L2: 40 B per ray (32 in + 8 out) + 112 B per node visited (7 x LDG.128) + 32 B per
                pre-tested leaf record + 64 B per exact test (4 vertices) -- served almost entirely by the caches, so the roof it is held
                against is the measured L2 bandwidth; the DRAM traffic of the ncu capture and the SURVEY 8(d) byte model (128 B nodes, 52 B
                primitives, against HBM) ride along.
      k_shade   max(HBM stream floor: 2 x 56 B path state + 20 B hit + 48 B shadow record per vertex, FP32 floor: ~900 flop-equivalents
                per Marschner vertex against 148 SMs x 128 lanes x 2 x SM clock)
      sort      k_ray_keys + 31-bit radix sort of (key, index) pairs: 40 B + 4 passes x 16 B + 8 B per entry, HBM
      raygen / splat   72 B written / 16 B read + 9 taps x 5 atomics x 4 B per path, HBM / L2 atomics"""
    rays = cnt['rays'] + cnt['shadow_rays_traced']
    nodes = cnt['nodes_visited'] + cnt['shadow_nodes_visited']; prims = cnt['prims_tested'] + cnt['shadow_prims_tested']
    exact = cnt['full_tests'] + cnt['shadow_full_tests']
    t_trace = max(prof['trace_ms'], 1e-9) * 1e-3
    loaded = 40.0 * rays + 112.0 * nodes + 32.0 * prims + 64.0 * exact
    survey = 40.0 * rays + 128.0 * nodes + 52.0 * prims
    ev = ncu_evidence() if scene == 'hair-curl' else None
    n_tr = max(1, prof['trace_launches'])
    trace = {'kernel': 'k_trace (BVH4 traversal of closest-hit + occlusion rays, fp32 pre-test, FP64 cylinder test)', 'bound': 'l2',
             'achieved': loaded / t_trace / 1e9, 'peak': l2_peak, 'unit': 'GB/s', 'frac': loaded / t_trace / 1e9 / l2_peak if l2_peak else None,
             'peak_source': 'L2 read bandwidth measured live (cudapath_measure_read_bandwidth, 32 MiB resident buffer)',
             'traffic': ev['dram_bytes'] if ev else None, 'traffic_note': ev['launch'] if ev else None,
             'bytes_per_launch': loaded / n_tr, 'avg_launch_ms': prof['trace_ms'] / n_tr, 'launches_per_step': n_tr,
             'mrays_per_s': rays / t_trace / 1e6, 'nodes_per_ray': nodes / max(1, rays), 'pretests_per_ray': prims / max(1, rays), 'exact_tests_per_ray': exact / max(1, rays),
             'loaded_bytes_per_ray': loaded / max(1, rays),
             'hbm_model_survey_8d': {'bytes_per_ray': survey / max(1, rays), 'achieved': survey / t_trace / 1e9, 'peak': hbm_peak, 'frac': survey / t_trace / 1e9 / hbm_peak,
                                     'note': 'SURVEY 8(d) byte model against HBM; these bytes are served by L1/L2, so this is NOT an HBM utilisation'},
             'ncu': ev,
             'binding': 'dependent-fetch latency and SIMT divergence (ncu: lanes per instruction, issue-slot utilisation, DRAM and L2 throughput all far below their peaks)'}
    verts = cnt['rays']                                  # one shaded vertex per closest-hit ray
    t_shade = max(prof['shade_ms'], 1e-9) * 1e-3
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12          # TFLOP/s, FMA = 2
    flops = 900.0 * verts; sbytes = (2 * 56 + 20 + 48) * verts
    f_fp = flops / t_shade / 1e12 / fp32_peak; f_hbm = sbytes / t_shade / 1e9 / hbm_peak
    shade = {'kernel': 'k_shade (frame, envmap sample, 2 x BSDF eval, BSDF sample, roulette, enqueue)', 'bound': 'fp32 issue' if f_fp >= f_hbm else 'hbm',
             'achieved': flops / t_shade / 1e12 if f_fp >= f_hbm else sbytes / t_shade / 1e9, 'peak': fp32_peak if f_fp >= f_hbm else hbm_peak,
             'unit': 'TFLOP/s (flop-equivalents)' if f_fp >= f_hbm else 'GB/s', 'frac': max(f_fp, f_hbm), 'frac_fp32': f_fp, 'frac_hbm': f_hbm,
             'vertices': verts, 'ms': prof['shade_ms'], 'flop_equivalents_per_vertex': 900, 'stream_bytes_per_vertex': 2 * 56 + 20 + 48}
    items = cnt['rays'] - cnt['paths'] + cnt['shadow_rays_traced']       # rays of bounces > 0 (camera rays are not sorted)
    t_sort = max(prof['sort_ms'], 1e-9) * 1e-3
    sb = 112.0 * items
    sort = {'kernel': 'k_ray_keys + cub::DeviceRadixSort (31-bit keys, library kernel)', 'bound': 'hbm', 'achieved': sb / t_sort / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': sb / t_sort / 1e9 / hbm_peak, 'entries': items, 'ms': prof['sort_ms'], 'bytes_per_entry': 112}
    t_rg = max(prof['raygen_ms'], 1e-9) * 1e-3; t_sp = max(prof['splat_ms'], 1e-9) * 1e-3
    raygen = {'kernel': 'k_raygen', 'bound': 'hbm', 'achieved': 72.0 * cnt['paths'] / t_rg / 1e9, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': 72.0 * cnt['paths'] / t_rg / 1e9 / hbm_peak, 'ms': prof['raygen_ms']}
    splat = {'kernel': 'k_splat (9 filter taps x 5 fp32 atomics per path)', 'bound': 'l2 atomics', 'achieved': (16.0 + 180.0) * cnt['paths'] / t_sp / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
             'frac': (16.0 + 180.0) * cnt['paths'] / t_sp / 1e9 / hbm_peak, 'ms': prof['splat_ms'], 'note': 'held against the HBM peak for lack of an atomic-throughput peak'}
    trace['stage_share_of_step'] = {k: prof[k + '_ms'] / max(prof['render_ms'], 1e-9) for k in ('trace', 'shade', 'sort', 'raygen', 'splat')}
    # sum of the per-launch stage times against the render time of the same (profiled) pass: the rest is launch gaps / host round trips
    trace['profiled_pass'] = {'render_ms': prof['render_ms'], 'stage_sum_ms': sum(prof[k + '_ms'] for k in ('trace', 'shade', 'sort', 'raygen', 'splat'))}
    head = {k: trace[k] for k in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'mrays_per_s', 'avg_launch_ms', 'launches_per_step')}
    trace['stages'] = [head, shade, sort, raygen, splat]
    trace['peak_sources'] = {'hbm': peak_src, 'l2': trace['peak_source'], 'fp32': '148 SMs x 128 lanes x 2 x %.0f MHz' % sm_mhz}
    return trace


def run_config5(args):
    """BASELINE.json configs[4]: 2^26 random Marschner eval+sample tuples and 2^26 random rays vs the furball BVH, per-stage rooflines.
    Inputs are generated on the device (torch, seed 0x5eed) and stay resident; CUDA events on the launching stream after warm-up."""
    import torch
    import cudapath
    n = 1 << args.log2n
    hbm_peak, sm_mhz, peak_src = load_peaks()
    torch.cuda.set_device(0)
    gen = torch.Generator(device='cuda'); gen.manual_seed(0x5eed)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = stream.cuda_stream
    warm = max(3, args.warmup)

    def sphere(m):
        v = torch.randn((m, 3), device='cuda', generator=gen)
        return (v / v.norm(dim=1, keepdim=True)).contiguous()

    def timeit(fn, reps):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    ctx = cudapath.scene_from_description('furball', scale=1.0)       # bsdf 0 = C3's Marschner block (ggx 0.2, IOR 1.55 / 1)
    kk = ctx.add_bsdf('kajiyakay', diffuseReflectance=HAIR_RGB, exponent=10.0)
    ctx.build(); build = ctx.stats()
    l2_peak = ctx.measure_read_bandwidth(32 << 20, 200)
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    sampler = ClockSampler(0); sampler.start(); t_begin = time.time()
    stages = []; launches = 0
    wi, wo = sphere(n), sphere(n); smp = torch.rand((n, 2), device='cuda', generator=gen)
    ev = torch.empty((n, 3), device='cuda'); pdf = torch.empty(n, device='cuda')
    swo = torch.empty((n, 3), device='cuda'); swt = torch.empty((n, 3), device='cuda'); sty = torch.empty(n, device='cuda', dtype=torch.int32)
    for name, bid, fl in (('marschner', 0, 900.0), ('kajiyakay', kk, 60.0)):
        t_eval = timeit(lambda: ctx.bsdf_eval_dev(bid, n, wi.data_ptr(), wo.data_ptr(), ev.data_ptr(), pdf.data_ptr(), sp), args.steps)
        t_smp = timeit(lambda: ctx.bsdf_sample_dev(bid, n, wi.data_ptr(), smp.data_ptr(), swo.data_ptr(), swt.data_ptr(), pdf.data_ptr(), sty.data_ptr(), sp), args.steps)
        launches += 2 * (args.steps + warm)
        t = (t_eval + t_smp) * 1e-3
        f_hbm = 72.0 * n / t / 1e9 / hbm_peak; f_fp = fl * n / t / 1e12 / fp32_peak
        stages.append({'stage': 'C5a %s eval+pdf+sample' % name, 'tuples': n, 'ms_eval_pdf': t_eval, 'ms_sample': t_smp, 'mtuples_per_s': n / t / 1e6,
                       'roofline': {'bound': 'fp32 issue' if f_fp >= f_hbm else 'hbm', 'frac': max(f_fp, f_hbm), 'frac_hbm_72B_per_tuple': f_hbm, 'frac_fp32_%d_flop_eq' % int(fl): f_fp,
                                    'floor_ms': max(72.0 * n / hbm_peak / 1e6, fl * n / fp32_peak / 1e9)}})
    del wi, wo, smp, ev, pdf, swo, swt, sty
    torch.cuda.empty_cache()
    aabb, bs = ctx.scene_bounds()
    c = torch.tensor(0.5 * (aabb[:3] + aabb[3:]), device='cuda'); r = float(0.5 * np.linalg.norm(aabb[3:] - aabb[:3]))   # bounding sphere of the fibers (kdbench.cpp:223-229)
    p1 = c + r * sphere(n); p2 = c + r * sphere(n)
    d = p2 - p1; d = (d / d.norm(dim=1, keepdim=True)).contiguous(); o = p1.contiguous()
    mint = torch.zeros(n, device='cuda'); maxt = torch.full((n,), float('inf'), device='cuda')
    sh = torch.empty(n, device='cuda', dtype=torch.int32); pr = torch.empty(n, device='cuda', dtype=torch.int32); tt = torch.empty(n, device='cuda')
    stats = torch.zeros(3, device='cuda', dtype=torch.int64)
    ray_stages = []

    def run_rays(label, o, d, mint, maxt):
        nonlocal launches
        for any_hit in (False, True):
            stats.zero_()
            ctx.intersect_dev(n, o.data_ptr(), d.data_ptr(), mint.data_ptr(), maxt.data_ptr(), sh.data_ptr(), pr.data_ptr(), tt.data_ptr(), any_hit, stats.data_ptr(), sp)
            torch.cuda.synchronize()
            nodes, prims, exact = int(stats[0]), int(stats[1]), int(stats[2])
            t = timeit(lambda: ctx.intersect_dev(n, o.data_ptr(), d.data_ptr(), mint.data_ptr(), maxt.data_ptr(), sh.data_ptr(), pr.data_ptr(), tt.data_ptr(), any_hit, 0, sp), args.steps)
            launches += args.steps + warm + 1
            loaded = 40.0 * n + 112.0 * nodes + 32.0 * prims + 64.0 * exact
            ray_stages.append({'stage': 'C5b %s %s' % (label, 'any-hit' if any_hit else 'closest-hit'), 'rays': n, 'ms': t, 'mrays_per_s': n / t / 1e3,
                               'hit_fraction': float((sh >= 0).float().mean()), 'nodes_per_ray': nodes / n, 'pretests_per_ray': prims / n, 'exact_tests_per_ray': exact / n,
                               'roofline': {'bound': 'l2', 'achieved': loaded / t / 1e6, 'peak': l2_peak, 'unit': 'GB/s', 'frac': loaded / t / 1e6 / l2_peak,
                                            'loaded_bytes_per_ray': loaded / n, 'frac_of_hbm_survey_model': (40.0 * n + 128.0 * nodes + 52.0 * prims) / t / 1e6 / hbm_peak}})

    run_rays('kdbench chords', o, d, mint, maxt)
    ctx.intersect_dev(n, o.data_ptr(), d.data_ptr(), mint.data_ptr(), maxt.data_ptr(), sh.data_ptr(), pr.data_ptr(), tt.data_ptr(), False, 0, sp)
    torch.cuda.synchronize()
    idx = torch.nonzero(sh >= 0).squeeze(1)
    idx = idx[torch.randint(0, len(idx), (n,), device='cuda', generator=gen)]
    o2 = (o[idx] + d[idx] * tt[idx].unsqueeze(1)).contiguous()
    nrm = sphere(n); u = sphere(n)
    d2 = nrm + u; d2 = (d2 / d2.norm(dim=1, keepdim=True).clamp_min(1e-6)).contiguous()     # cosine lobe about nrm
    run_rays('secondary-like (on fiber surfaces, cosine lobe, mint = Epsilon)', o2, d2, torch.full((n,), 1e-4, device='cuda'), maxt)
    clocks = sampler.finish(t_begin, time.time())
    head = ray_stages[0]
    line = {'metric': 'Mrays/s', 'value': head['mrays_per_s'], 'unit': 'Mrays/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': warm, 'ms_per_step': head['ms'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (fp64 cylinder test)', 'data': 'synthetic',
            'config': {'workload': 'BASELINE.json configs[4]: 2^%d random Marschner / Kajiya-Kay eval+pdf+sample tuples and 2^%d random rays vs the furball BVH (value = closest-hit kdbench chords)' % (args.log2n, args.log2n),
                       'l2_policy': 'inputs larger than L2 (2^%d-entry streams)' % args.log2n},
            'gpu_launches': launches, 'clocks': clocks, 'stages': stages + ray_stages,
            'roofline': dict(head['roofline'], kernel='k_intersect_batch<closest>', traffic=None, peak_source='L2 read bandwidth measured live'),
            'peaks': {'hbm_gbs': hbm_peak, 'l2_gbs': l2_peak, 'fp32_tflops': fp32_peak, 'source': peak_src},
            'bvh': {k: build[k] for k in ('segments', 'bvh_references', 'bvh_nodes', 'build_ms')}, 'e2e': None, 'cpu_baseline': None}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--scene', default='hair-curl', choices=['straight-hair', 'hair-curl', 'curly-hair', 'furball'])
    ap.add_argument('--scale', type=float, default=1.0, help='strand-count scale of the procedural generators')
    ap.add_argument('--spp', type=int, default=0, help='override samples per pixel (default: the config value)')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'], help='N > 1: split the fixed job (strong) or render spp per GPU (weak)')
    ap.add_argument('--shard', default='pixels', choices=['pixels', 'samples'], help='strong scaling: split the image by 32x32 pixel blocks or by sample ranges')
    ap.add_argument('--shard-test', default='', help='development: i/G renders only pixel shard i of G on this GPU')
    ap.add_argument('--config', type=int, default=2, help='5: the per-stage micro-benchmark of BASELINE.json configs[4]')
    ap.add_argument('--log2n', type=int, default=26, help='--config 5: log2 of the batch size')
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
    if args.config == 5:
        if rank == 0:
            run_config5(args)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    sc, shapes, env = scene_arrays(args.scene, args.scale)
    W, H, spp = sc['width'], sc['height'], sc['spp']
    shard = (0, 1)
    if args.scaling == 'weak':
        total_spp = spp * world; s_begin, s_end = rank * spp, (rank + 1) * spp          # every rank renders `spp` sample indices of a world*spp image
    elif args.shard == 'samples':
        total_spp = spp; s_begin, s_end = sample_range(spp, rank, world)               # the fixed image, sample indices split over the ranks
    else:
        total_spp = spp; s_begin, s_end = 0, spp; shard = (rank, world)                # the fixed image, 32x32 pixel blocks dealt out to the ranks
    if args.shard_test:                                                                # development: one GPU renders shard i of G of the image
        i, g = (int(v) for v in args.shard_test.split('/')); shard = (i, g)
    # camera paths this rank traces per step: picks the build effort of the scene (cudapath_set_job_size_hint)
    job_paths = W * H * (s_end - s_begin) // (shard[1] if shard[1] > 1 else 1)
    ctx = make_context(cudapath, sc, shapes, env, local, job_paths)
    ctx.set_pixel_shard(*shard)
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
        if s_end > s_begin:
            ctx.render_into(film.data_ptr(), total_spp, seed=seed, sample_begin=s_begin, sample_end=s_end, stream=stream.cuda_stream)
        if world > 1:
            cudapath.dist.reduce_film(film, dst=0)

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
    launches = 0; rays = 0; shadow = 0; waits = 0; step_render_ms = []
    e0.record(stream)
    for k in range(args.steps):
        step(1000 + k)
        if s_end > s_begin:
            st = ctx.stats(); step_render_ms.append(round(st['render_ms'], 2)); launches += st['kernel_launches'] + 1; rays += st['rays']; shadow += st['shadow_rays_traced']; waits += st['host_waits']      # rays actually traced (zero-contribution shadow rays are only counted)
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
    paths_per_step = W * H * total_spp
    if args.shard_test:
        paths_per_step = cnt['paths']
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
        host_film = torch.empty((H, W, 5), dtype=torch.float32).pin_memory()
        host_film_np = host_film.numpy()
        # W untimed warm-up runs (the first ones fill the caching allocator and find the GPU at idle clocks), then K timed runs
        times = []; phases = []
        n_warm = max(args.warmup, 0); n_timed = max(args.steps, 1)
        for k in range(n_warm + n_timed):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            c2 = make_context(cudapath, sc, pshapes, penv, local, job_paths)
            c2.set_pixel_shard(*shard)
            if args.wave:
                c2.set_options(wave_size=args.wave)
            t1 = time.perf_counter()
            c2.build()
            t2 = time.perf_counter()
            if world == 1:
                c2.render(total_spp, seed=2000 + k, sample_begin=s_begin, sample_end=s_end, out=host_film_np)        # host film out (page-locked)
            else:                                                                             # device film, NCCL reduce, read-back on rank 0
                film.zero_()
                if s_end > s_begin:
                    c2.render_into(film.data_ptr(), total_spp, seed=2000 + k, sample_begin=s_begin, sample_end=s_end, stream=stream.cuda_stream)
                cudapath.dist.reduce_film(film, dst=0)
                if rank == 0:
                    host_film.copy_(film, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if k >= n_warm:
                times.append(time.perf_counter() - t0)
                phases.append([round(t1 - t0, 4), round(t2 - t1, 4), round(time.perf_counter() - t2, 4), round(c2.stats()['render_ms'] * 1e-3, 4)])
            c2.close()
        te2e = torch.tensor([float(np.mean(times))], dtype=torch.float64, device='cuda')       # mean of the K timed runs
        if world > 1:
            dist.all_reduce(te2e, op=dist.ReduceOp.MAX)
        e2e = {'times_s': [round(t, 4) for t in times], 'phases_s(create+upload, build, render+reduce+readback, of which device render)': phases,
               'value': paths_per_step / float(te2e[0]) / 1e6, 'unit': 'Mpaths/s', 'h2d_bytes_per_step': int(h2d) * world, 'd2h_bytes_per_step': int(d2h),
               'warmup_runs': n_warm, 'includes': 'context creation, Marschner table build, geometry+envmap upload (every rank holds the scene), device BVH build, render, NCCL film reduce, film read-back'}

    if rank == 0:
        hbm_peak, sm_mhz, peak_src = load_peaks()
        l2_peak = ctx.measure_read_bandwidth(32 << 20, 200)
        roof = stage_rooflines(prof, cnt, hbm_peak, l2_peak, sm_mhz, peak_src, args.scene)
        cpu = None
        if not args.no_cpu and world == 1:
            cpu = run_reference(args, 0, 1)
        line = {'metric': 'Mpaths/s', 'value': value, 'unit': 'Mpaths/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
                'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32 (fp64 cylinder test)', 'data': 'synthetic',
                'config': workload_config(args, sc), 'mrays_per_s': mrays, 'rays_per_path': (rays + shadow) / args.steps / paths_per_step,
                'gpu_launches': int(launches), 'host_waits_per_step': waits / max(1, args.steps), 'step_render_ms': step_render_ms, 'clocks': clocks, 'e2e': e2e, 'roofline': roof, 'cpu_baseline': cpu,
                'build': {'segments': build['segments'], 'bvh_references': build['bvh_references'], 'bvh_nodes': build['bvh_nodes'], 'build_ms': build['build_ms'], 'build_ms_note': 'first build of the process (cold device allocator); the steady-state build is the second entry of e2e.phases_s'}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
