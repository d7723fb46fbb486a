#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2s}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
q() { echo "== $*" >> $out/${tag}_q.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_q.log; }
q bn254_g2 16,18,20
q bn254_g2 16,18,20 reduce_mode=2
q bn254_g2 20 acc_lockstep=0
q bn254_g2 20 reduce_roll=0
q bls12_381_g2 18,20
q bls12_381_g2 18,20 reduce_mode=2
q bls12_381_g2 20 acc_lockstep=0
q bls12_381 19,20,22
q bn254 16,20,24
cat $out/${tag}_pytest.log $out/${tag}_q.log
