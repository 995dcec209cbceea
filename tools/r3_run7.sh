#!/bin/bash
# full GPU suite + the bench line with generation 7 as the default decode path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3_tests7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3_tests7.log
timeout 900 python bench.py > gpurun_out/r3_bench7.json 2> gpurun_out/r3_bench7.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r3_bench7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3_bench7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])
for k,v in d['extras']['codecs'].items():
    if isinstance(v,dict): print(k, round(v['GBps'],1), round(v['ms'],2), round(v['frac'],4))
print(d['extras'].get('mixed_snappy_lz4_configs4',{}).get('GBps'), {k:(v.get('GBps') if isinstance(v,dict) else v) for k,v in d['extras'].get('real_corpus',{}).items() if 'decompress' in k})
PY
