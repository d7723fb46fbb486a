#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2z2}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python tools/quick_gpu.py bn254 12,14,16,18,20,21,22,23,24 2>&1 | grep "^2\^" > $out/${tag}_sizes_bn254.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1
cat $out/${tag}_pytest.log $out/${tag}_sizes_bn254.log; tail -2 $out/${tag}_smoke.log
python - <<'P'
import json
for l in open('gpurun_out/TAG_bench.json'.replace('TAG','r2z2')):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages_ms'], d['clocks'])
P
