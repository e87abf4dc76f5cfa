#!/bin/bash
# round-2 GPU run 9 (2 GPUs): the N-GPU film test, the new tests, bench.py under torchrun (strong scaling, both arms), the CLI on two devices
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
nvidia-smi -L | tee $out/g9_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or plastic_checkerboard or rectangle_and_mesh or teapot_scene or marschner_fixed_mode or native_cli" > $out/g9_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g9_pytest.log
tail -30 $out/g9_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $out/g9_bench_n2.json 2> $out/g9_bench_n2.err; echo "bench n2 rc=$?"
tail -3 $out/g9_bench_n2.err; python tools/dev/summ.py n2 < $out/g9_bench_n2.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $out/g9_bench_ref_n2.json 2> $out/g9_bench_ref_n2.err; echo "ref n2 rc=$?"
cut -c1-300 $out/g9_bench_ref_n2.json
timeout 600 python bench.py --steps 3 --warmup 3 > $out/g9_bench_n1.json 2> $out/g9_bench_n1.err; echo "bench n1 rc=$?"
python tools/dev/summ.py n1 < $out/g9_bench_n1.json | cut -c1-300
python - <<'PY' 2>&1 | tee $out/g9_cli.log
import os, subprocess, sys
sys.path.insert(0, '.')
import cudapath as cp
d = 'gpurun_out/g9_scene'; os.makedirs(d, exist_ok=True)
path = cp.scenes.write_scene('hair-curl', d, scale=0.05, overrides=dict(width=256, height=256, spp=16, maxDepth=16))
env = dict(os.environ, CUDAPATH_DATA_DIR=cp.DEFAULT_DATA_DIR)
for args in (['--gpus', '2'], ['-p', '2'], []):
    r = subprocess.run([cp.CLI_PATH] + args + ['-o', d + '/out%d.png' % len(args), path], capture_output=True, text=True, env=env)
    print(args, r.returncode, r.stdout.strip().splitlines()[-2:] if r.stdout else r.stderr[-300:])
PY
