# round-2 ncu captures of the three headline kernels (each after the same command passed without ncu)
Q="--steps 1 --warmup 3 --no-cpu --e2e-steps 1 --workloads none --dropin none --parity-points 2000 --parity-oracle-seconds 1"
cap() { # cfg points kernel-regex name
  python bench.py --config $1 --points $2 $Q > gpurun_out/plain_$4.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$3" -s 3 -c 1 -f -o gpurun_out/prof_$4 python bench.py --config $1 --points $2 $Q > gpurun_out/ncu_$4.log 2>&1; echo "$4 rc=$?"; }
cap 2 4000000 contract_mma cheb_mma_cfg2_r02
cap 4 2000000 contract_mma fh_mma_cfg4_r02
cap 3 20000000 mlip_cell mlip_cell_cfg3_r02
cap 5 500000 contract_mma cheb_mma_cfg5_r02
cap 1 1000000 cheb_slab cheb_slab_cfg1_r02
