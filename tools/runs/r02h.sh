# staggered warp groups of the DMMA engine: correctness, then the A/B sweep
python -m pytest tests/test_parity_gpu.py -m gpu -q -k "staggered or fh_fast or cheb_fast or full_size_properties or full_size_fh or 12pow6 or random_shapes" > gpurun_out/pytest_r02h.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_r02h.log
timeout 600 python tools/group_sweep.py > gpurun_out/group_sweep_r02.jsonl 2> gpurun_out/group_sweep_r02.err; echo "sweep rc=$?"; cut -c1-400 gpurun_out/group_sweep_r02.jsonl; tail -3 gpurun_out/group_sweep_r02.err
