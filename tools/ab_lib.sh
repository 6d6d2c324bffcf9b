#!/bin/bash
# tools/ab_lib.sh -- A/B of two builds of libchebpol_cuda on the GPU box (CHEBPOL_CUDA_LIB picks the library):
#   gpurun -- 'bash tools/ab_lib.sh tools/bin/libchebpol_cuda_nospline.so "1 2" > gpurun_out/ab.log 2>&1'
alt=$1; cfgs=${2:-1}
run() { CHEBPOL_CUDA_LIB=$1 python bench.py --config $2 --steps 20 --warmup 5 --no-cpu --e2e-steps 1 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3 cfg$2', d['value'], d['roofline']['kernel_ms'], d['roofline']['kernel'])"; }
for c in $cfgs; do for rep in 1 2 3; do run "" $c default; run "$PWD/$alt" $c alt; done; done
