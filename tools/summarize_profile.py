#!/usr/bin/env python3
"""Turns the ncu artefacts of tools/gpu_evidence.sh into the tracked summaries under profiles/.

  python tools/summarize_profile.py <tag> [--round r1]

Reads gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum launch list), gpurun_out/prof_<tag>_intersect.ncu-rep
(ncu --set full of one k_intersect launch; exported here with `ncu -i ... --page raw --csv`) and gpurun_out/bench_<tag>.json.
Writes profiles/<round>_launches_<tag>.csv (verbatim), profiles/<round>_kernel_shares_<tag>.md, profiles/<round>_k_intersect_ncu_<tag>.md
and profiles/k_intersect_traffic.json (read by bench.py for roofline.traffic).
"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, 'gpurun_out')
PROF = os.path.join(REPO, 'profiles')

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
    'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main():
    tag = sys.argv[1]
    rnd = sys.argv[sys.argv.index('--round') + 1] if '--round' in sys.argv else 'r1'
    os.makedirs(PROF, exist_ok=True)
    bench = None
    bp = os.path.join(OUT, 'bench_%s.json' % tag)
    if os.path.exists(bp):
        for line in open(bp):
            if line.startswith('{'):
                bench = json.loads(line)
        shutil.copy(bp, os.path.join(PROF, '%s_bench_%s.json' % (rnd, tag)))
    rp = os.path.join(OUT, 'bench_ref_%s.json' % tag)
    if os.path.exists(rp):
        shutil.copy(rp, os.path.join(PROF, '%s_bench_reference_%s.json' % (rnd, tag)))

    # ---- launch list -> per-kernel share
    lp = os.path.join(OUT, 'launches_%s.csv' % tag)
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(PROF, '%s_launches_%s.csv' % (rnd, tag)))
        rows = [r for r in csv.reader(l for l in open(lp) if l.startswith('"'))]
        hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            name = r[ki].split('(')[0].replace('void ', '')
            agg[name][0] += 1; agg[name][1] += float(r[vi].replace(',', '')) / 1e6
        total = sum(v[1] for v in agg.values())
        with open(os.path.join(PROF, '%s_kernel_shares_%s.md' % (rnd, tag)), 'w') as f:
            f.write('# Launch list summary (%s)\n\n' % tag)
            f.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e`\n')
            f.write('(cold-cache, serialised per-launch times; scene build + warm-up + profiled + counted + timed render; only the SHARE is comparable with bench.py)\n\n')
            f.write('| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
            for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write('| `%s` | %d | %.3f | %.1f %% |\n' % (name, n, ms, 100 * ms / total))
            render = {k: v for k, v in agg.items() if any(s in k for s in ('k_trace', 'k_ray_keys', 'k_intersect', 'k_shadow', 'k_shade', 'k_raygen', 'k_splat', 'k_sort_keys', 'DeviceRadixSort', 'k_apply'))}
            rt = sum(v[1] for v in render.values())
            f.write('\nRender-loop kernels only (%.2f ms): ' % rt + ', '.join('%s %.1f %%' % (k.split('::')[-1].split('<')[0], 100 * v[1] / rt) for k, v in sorted(render.items(), key=lambda kv: -kv[1][1])) + '\n')
            if bench:
                f.write('\nbench.py (same build, CUDA events, 64 spp): stage share of step = %s\n' % json.dumps(bench['roofline'].get('stage_share_of_step')))

    # ---- full capture of k_intersect
    kname = 'k_trace'
    rep = os.path.join(OUT, 'prof_%s_trace.ncu-rep' % tag)
    if not os.path.exists(rep):
        kname = 'k_intersect'; rep = os.path.join(OUT, 'prof_%s_intersect.ncu-rep' % tag)
    if os.path.exists(rep):
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        dram = to_bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + to_bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
        with open(os.path.join(PROF, '%s_%s_ncu_%s.md' % (rnd, kname, tag)), 'w') as f:
            f.write('# ncu --set full: one `%s` launch (%s)\n\n' % (kname, tag))
            f.write('Command: `ncu --set full --clock-control none --import-source on -k regex:%s -s 2 -c 1 python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e`\n' % kname)
            f.write('Kernel: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n' % d.get('Kernel Name'))
            for k in KEYS:
                if k in d:
                    f.write('| %s | %s | %s |\n' % (k, d[k], u[k]))
            f.write('\nDRAM traffic of this launch: %.1f MB (read + write).\n' % (dram / 1e6))
        def num(k):
            try: return float(d[k].replace(',', ''))
            except Exception: return None
        with open(os.path.join(PROF, 'k_trace_ncu.json' if kname == 'k_trace' else 'k_intersect_traffic.json'), 'w') as f:
            json.dump({'source': '%s_%s_ncu_%s.md' % (rnd, kname, tag), 'launch': 'third %s launch of bench.py --spp 8 --max-split 16 (secondary + shadow rays of one bounce, hair-curl; a launch of the 64-spp step is 8x this one)' % kname,
                       'dram_bytes': dram,
                       'lanes_per_instruction': num('smsp__thread_inst_executed_per_inst_executed.ratio'), 'issue_active_pct': num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                       'dram_throughput_pct': num('dram__throughput.avg.pct_of_peak_sustained_elapsed'), 'l2_throughput_pct': num('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
                       'l2_hit_pct': num('lts__t_sector_hit_rate.pct'), 'l1_hit_pct': num('l1tex__t_sector_hit_rate.pct'), 'warps_active_pct': num('sm__warps_active.avg.pct_of_peak_sustained_active'),
                       'registers': num('launch__registers_per_thread'), 'long_scoreboard_stall_per_issue': num('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
                       'duration_ms': float(d['gpu__time_duration.sum'].replace(',', '')) * ({'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[u['gpu__time_duration.sum']])}, f, indent=1)
    # ---- full captures of the other stages (tools/gpu_evidence_stages.sh): k_shade, k_shadow, the BSDF batch kernels of config 5a
    EXTRA = ['sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
             'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
             'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
             'sm__inst_executed_pipe_fp64.sum', 'sm__sass_thread_inst_executed_op_dfma_pred_on.sum', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum',
             'smsp__sass_thread_inst_executed_op_fp32_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fp64_pred_on.sum']
    for stage, cmd in (('shade', 'python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e  (-k regex:k_shade -s 2 -c 1)'), ('k_shade', 'python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e  (-k regex:k_shade -s 2 -c 1)'),
                       ('k_shadow', 'python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e  (-k regex:k_shadow -s 2 -c 1)'),
                       ('k_bsdf', 'python tools/microbench.py --log2n 24 --reps 1 --bsdf-only  (-k regex:k_bsdf_ -s 3 -c 2: Marschner eval+pdf, then sample)')):
        rep = os.path.join(OUT, 'prof_%s_%s.ncu-rep' % (tag, stage))
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
        hdr, units = rows[0], rows[1]
        with open(os.path.join(PROF, '%s_%s_ncu_%s.md' % (rnd, stage, tag)), 'w') as f:
            f.write('# ncu --set full: `%s` (%s)\n\nCommand: `ncu --set full --clock-control none --import-source on ... %s`\n' % (stage, tag, cmd))
            for vals in rows[2:]:
                d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
                f.write('\nKernel: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n' % d.get('Kernel Name'))
                for k in KEYS + EXTRA:
                    if k in d:
                        f.write('| %s | %s | %s |\n' % (k, d[k], u[k]))
    for extra in ('microbench_%s.jsonl' % tag, 'e2e_phases_%s.log' % tag, 'bench_config5_%s.json' % tag, 'pytest_gpu_%s.log' % tag):
        if os.path.exists(os.path.join(OUT, extra)):
            shutil.copy(os.path.join(OUT, extra), os.path.join(PROF, '%s_%s' % (rnd, extra)))
    print('profiles/ updated for', tag)


if __name__ == '__main__':
    main()
