# final 2-GPU sanity: the multi-GPU tests, then the driver's scaling command at N = 2 (both arms)
python -m pytest tests -m gpu -q -k "multi_gpu or sharded or stateless or cache or shard" > gpurun_out/pytest_r02_2gpu.log 2>&1; echo "2-gpu tests rc=$?"; tail -2 gpurun_out/pytest_r02_2gpu.log
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_2gpu_ref.json 2>/dev/null; echo "ref rc=$?"; grep -c '^{' gpurun_out/bench_r02_2gpu_ref.json
python - <<'PY'
import json
for l in open('gpurun_out/bench_r02_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'].get('all_ranks_ok'))
        for w in d['workloads']: print(w['cfg'], '%.4g'%w['value'], w['scaling'], 'frac %.3f'%w['roofline']['frac'], w['ms_per_step'], w['parity'].get('all_ranks_ok'))
        for e in d['e2e_dropin']: print({k:e.get(k) for k in ('cfg','n_gpus','points','value','host_gbs','vs_pinned_plan_e2e','error')})
PY
