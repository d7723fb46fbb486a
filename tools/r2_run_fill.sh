#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2f}
q() { echo "== $*" >> $out/${tag}_q.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep "^2\^" | cut -c1-100 >> $out/${tag}_q.log; }
for f in 0 48 96 192 384 768 1536; do q bn254_g2 16,20 level_fill=$f; done
for f in 48 96 192 384 768 1536; do q bls12_381 19,22 level_fill=$f; done
for f in 48 96 192 384 768; do q bls12_381_g2 20 level_fill=$f; done
cat $out/${tag}_q.log
