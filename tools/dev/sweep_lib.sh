#!/bin/bash
# usage: sweep_lib.sh <suffixes...> -- <bench args>   (runs bench.py once per alternative library build)
P=${GRAFT_REPO_ROOT:-/root/repo}/cs184-final-project-mitsuba0.5_b200
sufs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do sufs+=("$1"); shift; done; shift
for s in "${sufs[@]}"; do
  lib=$P/libcudapath$s.so
  CUDAPATH_LIB=$lib timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python tools/dev/summ.py "lib$s $*"
done
