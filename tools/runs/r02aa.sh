run() { CHEBPOL_CUDA_LIB=$1 python bench.py --config 5 --points 4000000 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 20000 --parity-oracle-seconds 3 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$2 cfg5', '%.4g' % d['value'], 'kernel_ms %.4f' % r['kernel_ms'], 'frac %.4f' % r['frac'], 'rel_err %.2e' % d['parity']['rel_err'])"; }
run "" default; run $PWD/tools/bin/libchebpol_cuda_ao.so oldA; run "" default; run $PWD/tools/bin/libchebpol_cuda_ao.so oldA
