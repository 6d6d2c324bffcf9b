export CHEBPOL_CUDA_TRACE=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 4 --steps 1 --warmup 3 --workloads none --dropin 2 --no-cpu --e2e-steps 1 --parity-points 2000 --parity-oracle-seconds 1 > gpurun_out/dropin4_trace.json 2> gpurun_out/dropin4_trace.err
echo rc=$?; grep "sharded call" gpurun_out/dropin4_trace.err | tail -22 | cut -c1-200; nproc
python - <<'PY'
import json
for l in open('gpurun_out/dropin4_trace.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for e in d['e2e_dropin']: print({k:e.get(k) for k in ('cfg','n_gpus','points','value','host_gbs','ms_per_step','vs_pinned_plan_e2e')})
PY
