#!/bin/bash
# round-2 GPU run 24 (2 GPUs): N-GPU film equality, strong-scaling bench under torchrun (both arms), CLI on two devices -- on the final library
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
nvidia-smi -L | tee $out/g24_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or pixel_shards or job_size or native_cli" > $out/g24_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g24_pytest.log
tail -5 $out/g24_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $out/g24_bench_n2.json 2> $out/g24_bench_n2.err; echo "bench n2 rc=$?"
tail -3 $out/g24_bench_n2.err; python tools/dev/summ.py n2 < $out/g24_bench_n2.json | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 > $out/g24_bench_n1.json 2> $out/g24_bench_n1.err; echo "bench n1 rc=$?"
python tools/dev/summ.py n1 < $out/g24_bench_n1.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --spp 512 --no-cpu > $out/g24_bench_n2_512spp.json 2> $out/g24_bench_n2_512spp.err; echo "bench n2 512spp rc=$?"
python tools/dev/summ.py n2-512spp < $out/g24_bench_n2_512spp.json | cut -c1-400
