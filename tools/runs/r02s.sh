run() { python bench.py --config $1 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 20000 --parity-oracle-seconds 3 $2 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ringdiv=$RD cfg$1', '%.4g' % d['value'], 'kernel_ms %.4f' % r['kernel_ms'], r['kernel'], 'frac %.4f' % r['frac'], 'rel_err %.2e' % d['parity']['rel_err'])"; }
for RD in 3 2; do export CHEBPOL_CUDA_MMA_RINGDIV=$RD; run 4; run 2; run 5 "--points 4000000"; done > gpurun_out/ringdiv_r02b.log 2>&1; cat gpurun_out/ringdiv_r02b.log
