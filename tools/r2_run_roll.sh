#!/bin/bash
# Code shape of the field multiplications (option `rolled`): parity suite, then stage timings
out=gpurun_out; mkdir -p $out; tag=${1:-r2q}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
q() { echo "== $*" >> $out/${tag}_roll.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_roll.log; }
for r in 12 13 14 7 11; do q bls12_381 19,22 rolled=$r; done
for r in 15 7 11 13; do q bn254 20,21,24 rolled=$r; done
for r in 12 13 14 7 11; do q bn254_g2 20 rolled=$r; done
for r in 12 13 7; do q bls12_381_g2 20 rolled=$r; done
cat $out/${tag}_pytest.log $out/${tag}_roll.log
