python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q -k "cheb or fh or full_size or general or random_shapes" > gpurun_out/pytest_r02k.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_r02k.log
bash tools/ab_lib.sh tools/bin/libchebpol_cuda_old.so "4 2 5" 2 > gpurun_out/ab_ginfo_r02.log 2>&1; cat gpurun_out/ab_ginfo_r02.log
