# 8 GPUs: drop-in legs cfg2 then cfg3 in one bench process (the sequence that exposed the staging-pool bug)
Q="--steps 2 --warmup 3 --workloads none --no-cpu --e2e-steps 5 --parity-points 2000 --parity-oracle-seconds 1"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 8 --dropin 2,3 $Q > gpurun_out/dropin8_both.json 2> gpurun_out/dropin8_both.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/dropin8_both.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['e2e']['value'])
        for e in d['e2e_dropin']: print({k:e.get(k) for k in ('cfg','n_gpus','points','value','host_gbs','ms_per_step','vs_pinned_plan_e2e','error')})
PY
