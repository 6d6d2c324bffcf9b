Q="--steps 1 --warmup 1 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 2000 --parity-oracle-seconds 1 --e2e-points 100000"
export CHEBPOL_CUDA_LIB=$PWD/tools/bin/libchebpol_cuda_tm.so
python bench.py --config 2 --points 4000000 $Q 2>&1 >/dev/null | grep -A14 "mma timing" | head -16
python bench.py --config 4 --points 4000000 $Q 2>&1 >/dev/null | grep -A14 "mma timing" | head -16
