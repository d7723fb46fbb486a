#!/bin/bash
out=gpurun_out
tag=${1:-r2m}
python tools/quick_gpu.py bn254 24 > $out/${tag}_q24.log 2>&1
python tools/quick_gpu.py bn254 24 segment=64 > $out/${tag}_q24_seg64.log 2>&1
python tools/quick_gpu.py bn254 24 segment=32 > $out/${tag}_q24_seg32.log 2>&1
python tools/quick_gpu.py bn254 21 segment=128 > $out/${tag}_q21_seg.log 2>&1
python tools/quick_gpu.py bn254 21 > $out/${tag}_q21.log 2>&1
python bench.py --steps 5 --no-extra > $out/${tag}_bench.json 2> $out/${tag}_bench.err
grep -h "^2\^" $out/${tag}_q24.log $out/${tag}_q24_seg64.log $out/${tag}_q24_seg32.log $out/${tag}_q21_seg.log $out/${tag}_q21.log | cut -c1-200
tail -c 300 $out/${tag}_bench.err; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('host_memory'), d['parity'])"
