#!/bin/bash
# round-2 GPU run 27 (1 GPU): where does the device idle between launches on a host that answers late?
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
nproc; uptime
timeout 300 python tools/dev/idle_probe.py 16 2>&1 | grep -v "queue reserve\|build:\|MemGetInfo: 0\.[0-4]" | tee $out/g27_idle_probe.log | tail -60
uptime
