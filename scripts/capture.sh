#!/bin/bash
# One-GPU evidence capture of the current build (called through gpurun): scripts/capture.sh <tag>
# Writes gpurun_out/<tag>_*: GPU test log, default bench line, reference arm, ncu launch list, ncu --set full of one
# C2 step and one C5 step (each profiled command first runs plain and must exit 0).
set -u
TAG=$1
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err; echo "bench rc=$?"
python bench.py --impl reference > $O/${TAG}_ref_n1.json 2> $O/${TAG}_ref_n1.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
python scripts/prof_step.py --warmup 2 --steps 1 > $O/${TAG}_prof_c2.json 2>&1 &&
ncu --set full --clock-control none --import-source on -s 10 -c 5 -f -o $O/${TAG}_c2 \
    python scripts/prof_step.py --warmup 2 --steps 1 > $O/${TAG}_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
python scripts/prof_step.py --workload c5 --warmup 2 --steps 1 > $O/${TAG}_prof_c5.json 2>&1 &&
ncu --set full --clock-control none --import-source on -s 10 -c 5 -f -o $O/${TAG}_c5 \
    python scripts/prof_step.py --workload c5 --warmup 2 --steps 1 > $O/${TAG}_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
ls -la $O/${TAG}_*
