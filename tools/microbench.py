#!/usr/bin/env python3
"""Per-stage micro-benchmarks of BASELINE.json config 5 (run on the GPU box):

  C5a  2^26 random BSDF tuples (wi, wo uniform on the sphere, 2 uniforms): eval+pdf and sample, Marschner block of C3 and the
       Kajiya-Kay block of C1.  Bound: max(72 B/tuple over HBM, FP issue) -- SURVEY.md section 8(d).
  C5b  2^26 rays against the furball BVH: kdbench-style chords of the bounding sphere (src/utils/kdbench.cpp:223-229) and
       "secondary-like" rays that start on fiber surfaces with cosine-distributed directions and mint = Epsilon.

Inputs are generated on the device (torch, seed 0x5eed) and stay resident; kernels are timed with CUDA events on the launching
stream after warm-up; one JSON line per stage.  Usage: python tools/microbench.py [--log2n 26] [--reps 5]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import cudapath

HAIR_RGB = (0.143016, 0.0156076, 1.80928e-005)


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    return (float(json.load(open(p))['hbm_gbs']), 'measured') if os.path.exists(p) else (6650.0, 'fallback')


def sphere(n, gen):
    v = torch.randn((n, 3), device='cuda', generator=gen)
    return (v / v.norm(dim=1, keepdim=True)).contiguous()


def timeit(fn, reps, stream):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--log2n', type=int, default=26)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--bsdf-only', action='store_true', help='stop after C5a (used for the ncu capture of the BSDF kernels)')
    args = ap.parse_args()
    n = 1 << args.log2n
    peak, peak_kind = peaks()
    gen = torch.Generator(device='cuda'); gen.manual_seed(0x5eed)
    stream = torch.cuda.Stream()          # a real (non-default) stream: handle 0 would make the library fall back to its own stream
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream

    # ---------------- C5a
    ctx = cudapath.scene_from_description('furball', scale=1.0)       # bsdf 0 = C3 Marschner block (ggx 0.2, IOR 1.55)
    kk = ctx.add_bsdf('kajiyakay', diffuseReflectance=HAIR_RGB, exponent=10.0)
    ctx.build()
    build = ctx.stats()
    wi, wo = sphere(n, gen), sphere(n, gen)
    smp = torch.rand((n, 2), device='cuda', generator=gen)
    ev = torch.empty((n, 3), device='cuda'); pdf = torch.empty(n, device='cuda')
    swo = torch.empty((n, 3), device='cuda'); swt = torch.empty((n, 3), device='cuda'); sty = torch.empty(n, device='cuda', dtype=torch.int32)
    for name, bid in (('marschner', 0), ('kajiyakay', kk)):
        t_eval, _ = timeit(lambda: ctx.bsdf_eval_dev(bid, n, wi.data_ptr(), wo.data_ptr(), ev.data_ptr(), pdf.data_ptr(), sp), args.reps, stream)
        t_smp, _ = timeit(lambda: ctx.bsdf_sample_dev(bid, n, wi.data_ptr(), smp.data_ptr(), swo.data_ptr(), swt.data_ptr(), pdf.data_ptr(), sty.data_ptr(), sp), args.reps, stream)
        by_eval = n * (24 + 12 + 4); by_smp = n * (12 + 8 + 12 + 12 + 4 + 4)
        for stage, t, by in (('eval+pdf', t_eval, by_eval), ('sample', t_smp, by_smp)):
            print(json.dumps({'stage': 'C5a bsdf %s %s' % (name, stage), 'tuples': n, 'ms': t, 'Mtuples_per_s': n / t / 1e3,
                              'roofline': {'bound': 'hbm', 'achieved': by / t / 1e6, 'peak': peak, 'unit': 'GB/s', 'frac': by / t / 1e6 / peak, 'peak_source': peak_kind,
                                           'bytes_per_tuple': by / n}}), flush=True)
    del wi, wo, smp, ev, pdf, swo, swt, sty
    if args.bsdf_only:
        return
    torch.cuda.empty_cache()

    # ---------------- C5b
    aabb, bs = ctx.scene_bounds()
    c = torch.tensor(0.5 * (aabb[:3] + aabb[3:]), device='cuda'); r = float(0.5 * np.linalg.norm(aabb[3:] - aabb[:3]))   # bounding sphere of the fibers
    p1 = c + r * sphere(n, gen); p2 = c + r * sphere(n, gen)
    d = p2 - p1; d = (d / d.norm(dim=1, keepdim=True)).contiguous(); o = p1.contiguous()
    mint = torch.zeros(n, device='cuda'); maxt = torch.full((n,), float('inf'), device='cuda')
    sh = torch.empty(n, device='cuda', dtype=torch.int32); pr = torch.empty(n, device='cuda', dtype=torch.int32); tt = torch.empty(n, device='cuda')
    stats = torch.zeros(2, device='cuda', dtype=torch.int64)

    def run_rays(label, o, d, mint, maxt):
        for any_hit in (False, True):
            stats.zero_()
            ctx.intersect_dev(n, o.data_ptr(), d.data_ptr(), mint.data_ptr(), maxt.data_ptr(), sh.data_ptr(), pr.data_ptr(), tt.data_ptr(), any_hit, stats.data_ptr(), sp)
            torch.cuda.synchronize()
            nodes, prims = int(stats[0]), int(stats[1])
            t, _ = timeit(lambda: ctx.intersect_dev(n, o.data_ptr(), d.data_ptr(), mint.data_ptr(), maxt.data_ptr(), sh.data_ptr(), pr.data_ptr(), tt.data_ptr(), any_hit, 0, sp),
                          args.reps, stream)
            by = 40.0 * n + 128.0 * nodes + 52.0 * prims
            print(json.dumps({'stage': 'C5b rays %s %s' % (label, 'any-hit' if any_hit else 'closest-hit'), 'rays': n, 'ms': t, 'Mrays_per_s': n / t / 1e3,
                              'hit_fraction': float((sh >= 0).float().mean()), 'nodes_per_ray': nodes / n, 'prims_per_ray': prims / n,
                              'roofline': {'bound': 'hbm', 'achieved': by / t / 1e6, 'peak': peak, 'unit': 'GB/s', 'frac': by / t / 1e6 / peak, 'peak_source': peak_kind,
                                           'bytes_per_ray': by / n},
                              'bvh': {k: build[k] for k in ('segments', 'bvh_references', 'bvh_nodes', 'build_ms')}}), flush=True)

    run_rays('chords', o, d, mint, maxt)
    # secondary-like: origins = closest hits of the chord batch, cosine-distributed directions about a random axis, mint = Epsilon
    ctx.intersect_dev(n, o.data_ptr(), d.data_ptr(), mint.data_ptr(), maxt.data_ptr(), sh.data_ptr(), pr.data_ptr(), tt.data_ptr(), False, 0, sp)
    torch.cuda.synchronize()
    hit = sh >= 0
    idx = torch.nonzero(hit).squeeze(1)
    idx = idx[torch.randint(0, len(idx), (n,), device='cuda', generator=gen)]
    o2 = (o[idx] + d[idx] * tt[idx].unsqueeze(1)).contiguous()
    nrm = sphere(n, gen); u = sphere(n, gen)
    d2 = nrm + u; d2 = (d2 / d2.norm(dim=1, keepdim=True).clamp_min(1e-6)).contiguous()     # cosine lobe about nrm
    mint2 = torch.full((n,), 1e-4, device='cuda')
    run_rays('secondary-like', o2, d2, mint2, maxt)


if __name__ == '__main__':
    main()
