# why is the pageable cfg3 path slow on the 8-GPU box?  host facts, then the drop-in leg alone and under torchrun
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA|MHz" | head -12; echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; nproc
python - <<'PY'
import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count(), 'OMP', os.environ.get('OMP_NUM_THREADS'))
PY
export CHEBPOL_CUDA_TRACE=1
Q="--steps 1 --warmup 3 --workloads none --no-cpu --e2e-steps 1 --parity-points 2000 --parity-oracle-seconds 1"
python bench.py --gpus 1 --config 3 --dropin 3 $Q > gpurun_out/dropin8_single.json 2> gpurun_out/dropin8_single.err; echo "single rc=$?"
grep "chebpol_cuda trace: device" gpurun_out/dropin8_single.err | tail -3 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --config 3 --dropin 3 $Q > gpurun_out/dropin8_torchrun.json 2> gpurun_out/dropin8_torchrun.err; echo "torchrun rc=$?"
grep "chebpol_cuda trace" gpurun_out/dropin8_torchrun.err | grep "device 0\|sharded call over" | tail -8 | cut -c1-300
top -b -n 1 | head -15
