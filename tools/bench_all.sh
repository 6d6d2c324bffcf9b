#!/bin/bash
# tools/bench_all.sh -- one bench line per named workload on one GPU (run on the GPU box):
#   gpurun --timeout 1500 -- 'bash tools/bench_all.sh > gpurun_out/bench_all.jsonl 2> gpurun_out/bench_all.err'
for c in 1 2 3 4 5 6 7 8; do
  python bench.py --config $c --steps 5 --warmup 3 --e2e-steps 2
done
