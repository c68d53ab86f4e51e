#!/bin/bash
# usage: tools/variants_env.sh "<tag>|<env assignments>|<nvcc flags>" ...   builds each variant on the GPU box and times the C2 scan
for v in "$@"; do
  tag="${v%%|*}"; rest="${v#*|}"; envs="${rest%%|*}"; flags="${rest#*|}"
  SC_EXTRA_NVCC="$flags" python surfcascade_b200/build.py --force > /dev/null 2>gpurun_out/build_$tag.err || { echo "{\"tag\": \"$tag\", \"build\": \"failed\"}"; continue; }
  env $envs python tools/quick_stage0.py 32 "$tag" 2>>gpurun_out/build_$tag.err
done
SC_EXTRA_NVCC="" python surfcascade_b200/build.py --force > /dev/null 2>&1
