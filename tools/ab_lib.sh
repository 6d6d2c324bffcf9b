#!/bin/bash
# tools/ab_lib.sh -- A/B of two builds of libchebpol_cuda on the GPU box (CHEBPOL_CUDA_LIB picks the library):
#   make -C chebpol_b200/csrc BUILD=build_alt OUT=../../tools/bin/libchebpol_cuda_alt.so EXTRA=-DSOMETHING
#   gpurun -- 'bash tools/ab_lib.sh tools/bin/libchebpol_cuda_alt.so "4 2" > gpurun_out/ab.log 2>&1'
alt=$1; cfgs=${2:-1}; reps=${3:-2}
run() { CHEBPOL_CUDA_LIB=$1 python bench.py --config $2 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none \
        --parity-points 20000 --parity-oracle-seconds 3 $4 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$3 cfg$2', '%.4g' % d['value'], 'kernel_ms %.4f' % r['kernel_ms'], r['kernel'], 'frac %.4f' % r['frac'], 'rel_err %.2e' % d['parity']['rel_err'])"; }
for c in $cfgs; do
  extra=""; [ "$c" = "5" ] && extra="--points 4000000"
  for rep in $(seq $reps); do run "" $c default "$extra"; run "$PWD/$alt" $c alt "$extra"; done
done
