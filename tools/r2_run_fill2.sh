#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2g}
q() { echo "== $*" >> $out/${tag}_q.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep "^2\^" | cut -c1-150 >> $out/${tag}_q.log; }
for f in 24 48 96 192 384 768 1536; do q bn254 14,16,18,19,20,21,22,23,24 level_fill=$f; done
for f in 24 48 96 192 384 768; do q bls12_381 14,16,18,20,21 level_fill=$f; done
cat $out/${tag}_q.log
