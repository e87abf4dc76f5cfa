#!/bin/bash
# usage: ab.sh <rounds> <variants...> -- <bench args>   interleaved A/B: every variant is benchmarked <rounds> times, round-robin
# a variant is "base", a library suffix ("_pf" -> libcudapath_pf.so), "env:NAME=VALUE[,NAME2=VALUE2]" (base library with environment variables) or "_pf+NAME=VALUE[,...]" (both)
P=${GRAFT_REPO_ROOT:-/root/repo}/cs184-final-project-mitsuba0.5_b200
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
rounds=$1; shift
sufs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do sufs+=("$1"); shift; done; shift
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e "$@" > /dev/null 2>&1   # warm the box up
for r in $(seq $rounds); do for s in "${sufs[@]}"; do
  lib=$P/libcudapath.so; envs=""
  case "$s" in base) ;; env:*) envs="${s#env:}";; *+*) lib=$P/libcudapath${s%%+*}.so; envs="${s#*+}";; *) lib=$P/libcudapath$s.so;; esac
  envs="${envs//,/ }"
  env $envs CUDAPATH_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python tools/dev/summ.py "r$r $s $*" | cut -c1-330
done; done
