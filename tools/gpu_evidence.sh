#!/bin/bash
# Evidence run on the GPU box (one GPU): parity tests, bench (both arms), config-5 microbench, ncu launch list, one ncu --set full capture of k_trace.
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_evidence.sh <tag>'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -3 $out/pytest_gpu_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python tools/dev/summ.py default < $out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"
cut -c1-400 $out/bench_ref_$tag.json
timeout 600 python bench.py --config 5 --steps 3 > $out/bench_config5_$tag.json 2> $out/bench_config5_$tag.err; echo "config5 rc=$?"
# launch list (cold-cache, serialised): same command at 8 spp so it stays short
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_$tag.csv \
    python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
# full capture of one mid-render k_trace launch (second bounce of the first wave)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -o $out/prof_${tag}_trace \
    python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 1 -o $out/prof_${tag}_shade \
    python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_full_shade_$tag.log 2>&1; echo "ncu full shade rc=$?"
