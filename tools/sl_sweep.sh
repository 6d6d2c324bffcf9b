#!/bin/bash
# tools/sl_sweep.sh -- sl_cell tuning sweep on the GPU box: batch / min-blocks variants, then cell-grid resolutions.
#   gpurun --timeout 900 -- 'bash tools/sl_sweep.sh > gpurun_out/sl_sweep.log 2>&1'
#   SL_CELLS="200 256" SL_VARIANTS="" bash tools/sl_sweep.sh      # only some resolutions
run() { python bench.py --config 8 --variant "$1" --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['value'], d['roofline']['kernel_ms'], d['parity']['max_abs_err'])"; }
for v in ${SL_VARIANTS-1 2 4 8 802 804}; do run $v "variant=$v"; done
for c in ${SL_CELLS-41 64 128 160}; do CHEBPOL_CUDA_SL_CELLS=$c run ${SL_V:-804} "cells=$c"; done
