# 64-byte L2 fills on the quad and brick gathers: timings, then DRAM bytes of one launch each
python -m pytest tests -m gpu -x -q -k "mlip" > gpurun_out/pytest_r02o_mlip.log 2>&1; echo "mlip tests rc=$?"; tail -2 gpurun_out/pytest_r02o_mlip.log
(python tools/mlip_variants.py --points 1e8 --variants 3,1; CHEBPOL_CUDA_CELL_M=3 python tools/mlip_variants.py --points 1e8 --variants 2; CHEBPOL_CUDA_CELL_M=2 python tools/mlip_variants.py --points 1e8 --variants 2; python tools/mlip_variants.py --points 1e8 --variants 2; CHEBPOL_CUDA_CELL_MAX_GB=40 CHEBPOL_CUDA_CELL_M=3 python tools/mlip_variants.py --shape 48,48,48,48,48 --points 1e8 --variants 2,3) > gpurun_out/mlip_variants_r02o.jsonl 2> gpurun_out/mlip_variants_r02o.err
cut -c1-330 gpurun_out/mlip_variants_r02o.jsonl; tail -2 gpurun_out/mlip_variants_r02o.err
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv"
ncu $M -k regex:mlip_quad_kernel -s 3 -c 1 --log-file gpurun_out/dram_quad_r02o.csv python tools/mlip_variants.py --points 3e7 --variants 3 --steps 1 > /dev/null 2>&1
CHEBPOL_CUDA_CELL_M=3 ncu $M -k regex:mlip_cell_kernel -s 3 -c 1 --log-file gpurun_out/dram_brick3_r02o.csv python tools/mlip_variants.py --points 3e7 --variants 2 --steps 1 > /dev/null 2>&1
grep -h "dram__bytes\|gpu__time" gpurun_out/dram_quad_r02o.csv gpurun_out/dram_brick3_r02o.csv | cut -d, -f5,13- | cut -c1-160
