Q="--steps 2 --warmup 1 --sub-seconds 0.02 --no-cpu --dropin none --parity-points 2000 --parity-oracle-seconds 1 --e2e-steps 1"
python bench.py $Q > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py $Q > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
