#!/bin/bash
# Round-2 A/B runs on one B200 (called through gpurun): autocorrelation MMA-loop variants and host-pipeline group sizes.
# Every line of gpurun_out/r2_ab.jsonl is one bench.py JSON line prefixed by the variant tag.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab.jsonl
: > $OUT
run() {  # tag, env assignments..., -- bench args
  local tag=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  local line
  line=$(env "${envs[@]}" python bench.py --no-extras --no-cpu-baseline "$@" 2>>gpurun_out/r2_ab.err | tail -1)
  echo "{\"tag\": \"$tag\", \"line\": $line}" >> $OUT
}
for loop in legacy alt ring; do
  run "c2_$loop" NSF_AC_LOOP=$loop -- --workload c2 --steps 20 --warmup 5
  run "c5_$loop" NSF_AC_LOOP=$loop -- --workload c5 --steps 10 --warmup 3
done
for g in 8 16 32; do
  for f in 2 6; do
    run "c2_group${g}_first${f}" NSF_GROUP_MI=$g NSF_FIRST_MI=$f -- --workload c2 --steps 20 --warmup 5
  done
done
