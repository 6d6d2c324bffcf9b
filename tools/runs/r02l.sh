python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q -k "cheb or fh or full_size or general or random_shapes" > gpurun_out/pytest_r02l.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_r02l.log
run() { python bench.py --config $1 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 20000 --parity-oracle-seconds 3 $2 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg$1', '%.4g' % d['value'], 'kernel_ms %.4f' % r['kernel_ms'], r['kernel'], 'frac %.4f' % r['frac'], 'rel_err %.2e' % d['parity']['rel_err'])"; }
(run 2; run 5 "--points 4000000"; run 4) > gpurun_out/pairrows_r02.log 2>&1; cat gpurun_out/pairrows_r02.log
