# round-2 evidence run on one B200: tests, bench (both arms), crossover table, ncu launch list and captures
python -m pytest tests -m gpu -q > gpurun_out/pytest_r02g.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/pytest_r02g.log
python bench.py > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02g.err
python bench.py --impl reference > gpurun_out/bench_r02g_ref.json 2>/dev/null; echo "ref rc=$?"
python tools/crossover_table.py > gpurun_out/crossover_r02.jsonl 2> gpurun_out/crossover_r02.err; echo "crossover rc=$?"; tail -2 gpurun_out/crossover_r02.err
Q="--steps 2 --warmup 1 --sub-seconds 0.02 --no-cpu --dropin none --parity-points 2000 --parity-oracle-seconds 1 --e2e-steps 1"
python bench.py $Q > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py $Q > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/mlip_variants.py --points 3e7 --variants 3 --steps 1 > gpurun_out/plain_q.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlip_quad_kernel -s 3 -c 1 -o gpurun_out/prof_mlip_quad_cfg3_r02 python tools/mlip_variants.py --points 3e7 --variants 3 --steps 1 > gpurun_out/ncu_q.log 2>&1; echo "quad ncu rc=$?"
CHEBPOL_CUDA_CELL_M=3 python tools/mlip_variants.py --points 3e7 --variants 2 --steps 1 > gpurun_out/plain_b.log 2>&1 && CHEBPOL_CUDA_CELL_M=3 ncu --set full --clock-control none --import-source on -k regex:mlip_cell_kernel -s 3 -c 1 -o gpurun_out/prof_mlip_brick3_cfg3_r02 python tools/mlip_variants.py --points 3e7 --variants 2 --steps 1 > gpurun_out/ncu_b.log 2>&1; echo "brick ncu rc=$?"
