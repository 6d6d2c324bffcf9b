python -m pytest tests -m gpu -q -k "mlip or stateless or chunk or device_resident or empty or multi_gpu or cache" > gpurun_out/pytest_r02v.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_r02v.log
python bench.py --config 4 --workloads 1,3 --dropin 3 --no-cpu --sub-seconds 0.5 --parity-points 20000 --parity-oracle-seconds 3 > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02v.json'))
print('cfg4 value %.4g e2e %.4g ratio %.3f'%(d['value'], d['e2e']['value'], d['e2e']['value']/d['value']))
for w in d['workloads']: print('cfg',w['cfg'],'value %.4g e2e %.4g'%(w['value'], w['e2e']['value']))
for e in d['e2e_dropin']: print({k:e.get(k) for k in ('cfg','n_gpus','value','host_gbs','ms_per_step')})
PY
