t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r02_4gpu.json 2> gpurun_out/bench_r02_4gpu.err
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_4gpu.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'].get('all_ranks_ok'))
        for w in d['workloads']: print(w['cfg'], '%.4g'%w['value'], w['scaling'], 'frac %.3f'%w['roofline']['frac'], w['ms_per_step'], w['parity'].get('all_ranks_ok'))
        for e in d['e2e_dropin']: print({k:e.get(k) for k in ('cfg','n_gpus','points','value','host_gbs','vs_pinned_plan_e2e','error')})
PY
