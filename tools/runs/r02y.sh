run() { python bench.py --config $1 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 20000 --parity-oracle-seconds 3 $2 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg$1', '%.4g' % d['value'], 'kernel_ms %.4f' % r['kernel_ms'], r['kernel'], 'frac %.4f' % r['frac'], 'pipe %.4f' % r.get('pipe_frac',0), 'rel_err %.2e' % d['parity']['rel_err'])"; }
(run 5 "--points 4000000"; run 2; run 4) 2>&1
