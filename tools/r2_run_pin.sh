#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2z3}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_pytest.log
q() { echo "== $*" >> $out/${tag}_q.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_q.log; }
q bn254 16,20,21,22,23,24
q bn254 24 dist=witness
q bls12_381 19,22
q bn254_g2 20
cat $out/${tag}_pytest.log $out/${tag}_q.log
