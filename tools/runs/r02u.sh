python -m pytest tests -m gpu -q > gpurun_out/pytest_r02z.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/pytest_r02z.log
python bench.py > gpurun_out/bench_r02z.json 2> gpurun_out/bench_r02z.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02z.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02z.json'))
r=d['roofline']; print('headline', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'frac %.4f'%r['frac'], 'pipe', r.get('pipe_frac'), 'cpu', d['cpu_baseline']['value'])
for w in d['workloads']:
    r=w['roofline']; print(w['cfg'], '%.4g'%w['value'], 'e2e %.4g'%w['e2e']['value'], 'frac %.4f'%r['frac'], 'best', r.get('frac_best'), 'pipe', r.get('pipe_frac'), w['clocks']['reasons'], w['parity']['worst_rel_err_over_ranks'])
for e in d['e2e_dropin']: print({k:e.get(k) for k in ('cfg','n_gpus','value','host_gbs','vs_pinned_plan_e2e')})
PY
python bench.py --impl reference > gpurun_out/bench_r02z_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_r02z_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
