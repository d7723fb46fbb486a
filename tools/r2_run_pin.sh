#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2z5}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_pytest.log
q() { echo "== $*" >> $out/${tag}_q.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep "^2\^" | cut -c1-118 >> $out/${tag}_q.log; }
q bn254 12,14,16,17,18,19,20,21,22,23,24
q bls12_381 16,18,19,20,22
q bn254_g2 16,18,20
q bls12_381_g2 16,18,20
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cat $out/${tag}_pytest.log $out/${tag}_q.log
