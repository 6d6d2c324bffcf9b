#!/bin/bash
# tools/profile.sh -- the ncu recipe behind profiles/*.txt.  Run ON the GPU box, ONE capture per
# call (ncu replays each kernel ~40 times), always after the same command has exited 0 without ncu:
#   gpurun --timeout 1200 -- 'bash tools/profile.sh full 2 4000000 contract_mma cheb_mma_cfg2_r01'
#   gpurun --timeout 1200 -- 'bash tools/profile.sh launches 2 4000000 - cfg2_r01'
# then, back in the container:
#   python tools/ncu_summary.py gpurun_out/prof_<name>.ncu-rep profiles/ncu_<name>.txt <points>
set -euo pipefail
mode=$1; cfg=$2; pts=$3; kernel=$4; name=$5
cmd="python bench.py --config $cfg --points $pts --steps 1 --warmup 3 --no-cpu --e2e-steps 1"
mkdir -p gpurun_out
$cmd > gpurun_out/b_small.log 2>&1   # must pass on its own first; a number printed under ncu is never a bench value
if [ "$mode" = full ]; then
  ncu --set full --clock-control none --import-source on -k "regex:$kernel" -s 3 -c 1 -f \
      -o "gpurun_out/prof_$name" $cmd > gpurun_out/ncu_f.log 2>&1
else
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file "gpurun_out/launches_$name.csv" $cmd > gpurun_out/ncu_l.log 2>&1
fi
echo "ncu rc=$?"
