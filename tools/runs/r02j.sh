# drop-in leg at 2 GPUs with the host-path trace: where does the one-process sharded call lose time?
export CHEBPOL_CUDA_TRACE=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 3 --workloads none --dropin 2,3 --no-cpu --e2e-steps 1 --parity-points 2000 --parity-oracle-seconds 1 > gpurun_out/dropin2_trace.json 2> gpurun_out/dropin2_trace.err
echo rc=$?; grep "chebpol_cuda trace" gpurun_out/dropin2_trace.err | tail -30
python - <<'PY'
import json
for l in open('gpurun_out/dropin2_trace.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for e in d['e2e_dropin']: print({k:e.get(k) for k in ('cfg','n_gpus','points','value','host_gbs','ms_per_step','vs_pinned_plan_e2e','error')})
PY
nproc; free -g | head -2
python -m pytest tests/test_parity_gpu.py -m gpu -q -k "multi_gpu or stateless or cache" 2>&1 | tail -3
