# FH/DMMA engine experiments (B prefetch across steps, ring geometry) and the gather-granularity microbenchmark
bash tools/ab_lib.sh tools/bin/libchebpol_cuda_e1.so "4 2 5" 2 > gpurun_out/ab_e1_r02.log 2>&1; cat gpurun_out/ab_e1_r02.log
run() { python bench.py --config $1 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 20000 --parity-oracle-seconds 3 $2 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ringdiv=$RD cfg$1', '%.4g' % d['value'], 'kernel_ms %.4f' % r['kernel_ms'], r['kernel'], 'frac %.4f' % r['frac'], 'rel_err %.2e' % d['parity']['rel_err'])"; }
for RD in 2 4 6; do export CHEBPOL_CUDA_MMA_RINGDIV=$RD; run 4; run 2; done > gpurun_out/ringdiv_r02.log 2>&1; unset CHEBPOL_CUDA_MMA_RINGDIV; cat gpurun_out/ringdiv_r02.log
tools/bin/gather_granule 4096 > gpurun_out/gather_granule_r02.jsonl 2>&1; cat gpurun_out/gather_granule_r02.jsonl
ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gather_granule_ncu_r02.csv tools/bin/gather_granule 4096 > /dev/null 2>&1; echo "ncu rc=$?"
