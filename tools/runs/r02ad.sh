run() { CHEBPOL_CUDA_LIB=$1 python bench.py --config $3 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 20000 --parity-oracle-seconds 3 $4 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$2 cfg$3', '%.4g' % d['value'], 'frac %.4f' % r['frac'], 'rel_err %.2e' % d['parity']['rel_err'])"; }
for v in default T1 T2 T12; do lib=""; [ $v != default ] && lib=$PWD/tools/bin/libchebpol_cuda_$v.so; run "$lib" $v 4; run "$lib" $v 2; run "$lib" $v 5 "--points 4000000"; done
