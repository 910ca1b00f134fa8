#!/bin/bash
# Usage: profiles/store_records.sh <old-tag> <new-tag> <bench.json> <ref.json>   (run from the repo root after a validation run)
set -e
OLD=$1; NEW=$2; BENCH=$3; REF=$4
cd "$(dirname "$0")/.."
git rm -q --cached profiles/r1/step_f32_16M_${OLD}_pdl.ncu_summary.txt profiles/r1/step_f32_1M_${OLD}_pdl.ncu_summary.txt profiles/r1/launches_bench_steps40_${OLD}.csv profiles/r1/launches_bench_steps40_${OLD}.summary.txt profiles/r1/bench_${OLD}_final.json profiles/r1/bench_${OLD}_reference_arm.json 2>/dev/null || true
rm -f profiles/r1/step_f32_16M_${OLD}_pdl.ncu_summary.txt profiles/r1/step_f32_1M_${OLD}_pdl.ncu_summary.txt profiles/r1/launches_bench_steps40_${OLD}.csv profiles/r1/launches_bench_steps40_${OLD}.summary.txt profiles/r1/bench_${OLD}_final.json profiles/r1/bench_${OLD}_reference_arm.json
python profiles/ncu_summary.py gpurun_out/step_f32_16M_${NEW}.ncu-rep > profiles/r1/step_f32_16M_${NEW}_pdl.ncu_summary.txt 2>&1
python profiles/ncu_summary.py gpurun_out/step_f32_1M_${NEW}.ncu-rep > profiles/r1/step_f32_1M_${NEW}_pdl.ncu_summary.txt 2>&1
cp gpurun_out/launches_${NEW}.csv profiles/r1/launches_bench_steps40_${NEW}.csv
{ echo "ncu --metrics gpu__time_duration.sum --clock-control none -c 700 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline"
  echo "(the same command exited 0 without ncu immediately before; per-launch times are cold-cache and serialised: compare SHARES)"; echo
  python profiles/launch_summary.py gpurun_out/launches_${NEW}.csv; echo
  echo "Inside the timed region bench.py launches only cartpole_step_kernel<float, true, false> (one per step): share of the step = 100 %."
  echo "The other launches are set-up (action buffers, initial reset, torch fills) and the e2e leg's copies."; } > profiles/r1/launches_bench_steps40_${NEW}.summary.txt
cp "$BENCH" profiles/r1/bench_${NEW}_final.json; cp "$REF" profiles/r1/bench_${NEW}_reference_arm.json
sed -i "s/bench_${OLD}_final/bench_${NEW}_final/g; s/bench_${OLD}_reference_arm/bench_${NEW}_reference_arm/g; s/launches_bench_steps40_${OLD}/launches_bench_steps40_${NEW}/g; s/step_f32_16M_${OLD}_pdl/step_f32_16M_${NEW}_pdl/g; s/step_f32_1M_${OLD}_pdl/step_f32_1M_${NEW}_pdl/g" profiles/r1/README.md DESIGN.md README.md profiles/traffic.json
grep -E "duration|dram__bytes" profiles/r1/step_f32_16M_${NEW}_pdl.ncu_summary.txt
