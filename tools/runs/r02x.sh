python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q -k "cheb or fh or full_size or general or random_shapes or evalongrid" > gpurun_out/pytest_r02x.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_r02x.log
run() { python bench.py --config $1 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 20000 --parity-oracle-seconds 3 $2 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg$1', '%.4g' % d['value'], 'kernel_ms %.4f' % r['kernel_ms'], r['kernel'], 'frac %.4f' % r['frac'], 'pipe %.4f' % r.get('pipe_frac',0), 'rel_err %.2e' % d['parity']['rel_err'])"; }
(run 4; run 2; run 5 "--points 4000000") > gpurun_out/local_tables_r02.log 2>&1; cat gpurun_out/local_tables_r02.log
Q="--steps 1 --warmup 1 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 2000 --parity-oracle-seconds 1 --e2e-points 100000"
export CHEBPOL_CUDA_LIB=$PWD/tools/bin/libchebpol_cuda_tm.so
python bench.py --config 2 --points 4000000 $Q 2>&1 >/dev/null | grep -A12 "mma timing" | head -13
python bench.py --config 4 --points 4000000 $Q 2>&1 >/dev/null | grep -A12 "mma timing" | head -13
