#!/bin/bash
# Round-2 evidence run (one gpurun call, one GPU; `bash profiles/r2_capture.sh`): plain bench (both arms), the ncu launch list of the same bench command,
# and one `ncu --set full` capture per kernel family.  Every ncu command is preceded by the same command run plainly.
set -x
O=gpurun_out/r2/final; mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline > $O/plain_launchlist.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
      python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline > $O/ncu_launchlist.log 2>&1
cap() {  # name, kernel regex, skip, command...
  name=$1; regex=$2; skip=$3; shift 3
  "$@" > $O/plain_$name.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/$name -f "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  # the reports are 10-20 MB each and gpurun brings back at most 64 MiB: summarise here, keep the text
  python profiles/ncu_summary.py /tmp/$name.ncu-rep > $O/$name.ncu_summary.txt 2>&1
}
cap step_f32_16M cartpole_step_kernel 34 python profiles/drive.py step --n 16777216 --iters 6
cap step_lean_f32_16M cartpole_step_kernel 34 python profiles/drive.py step --n 16777216 --iters 6 --lean
cap step_f32_1M_tile cartpole_step_kernel 34 python profiles/drive.py step --n 1048576 --iters 6
cap sample_f32_uniform_16M dr_sample 1 python profiles/drive.py sample --n 16777216 --dr uniform --iters 3
cap sample_f32_gaussian_16M dr_sample 1 python profiles/drive.py sample --n 16777216 --dr gaussian --iters 3
cap sample_f32_truncnorm_16M dr_sample 1 python profiles/drive.py sample --n 16777216 --dr truncnorm --iters 3
cap fullgaussian_tc_4M fullgaussian_tc 1 python profiles/exp/drive_fullgaussian.py
cap rollout_pair_f32_4M_K100 rollout_pair 0 python profiles/drive.py rollout --n 4194304 --iters 1 --K 100
cap rollout_pair_resetheavy_4M_K300 rollout_pair 0 python profiles/drive.py rollout --n 4194304 --iters 1 --K 300 --policy resetheavy
cap rollout_random_f32_4M_K100 cartpole_rollout_kernel 0 python profiles/drive.py rollout --n 4194304 --iters 1 --K 100 --policy random
ls -la $O
