(python tools/ring_sweep.py; CHEBPOL_CUDA_MMA_RINGDIV=2 python tools/ring_sweep.py; CHEBPOL_CUDA_MMA_RINGDIV=2 CHEBPOL_CUDA_MMA_SLAB_KB=128 python tools/ring_sweep.py) > gpurun_out/ring_sweep_r02.jsonl 2> gpurun_out/ring_sweep_r02.err; tail -3 gpurun_out/ring_sweep_r02.err
python - <<'PY'
import json,collections
t=collections.defaultdict(dict)
for l in open('gpurun_out/ring_sweep_r02.jsonl'):
    d=json.loads(l); t[(d['kind'],tuple(d['dims']))][(d['ringdiv'],d['slab_kb'])]=(d['frac_fp64_peak'],d['checksum'])
for k,v in t.items(): print(k, ' '.join('%s/%s: %.4f'%(a,b,f) for (a,b),(f,c) in v.items()), 'checksums equal' if len({c for f,c in v.values()})==1 else 'CHECKSUMS DIFFER')
PY
