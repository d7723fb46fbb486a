#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2v}
q() { echo "== $*" >> $out/${tag}_arena.log; env "$@" timeout 300 python tools/quick_gpu.py bn254 21,24 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_arena.log; }
q TACHYON_B200_ARENA=1
q TACHYON_B200_ARENA=1 TACHYON_B200_PAD_MB=62
q TACHYON_B200_ARENA=1 TACHYON_B200_PAD_MB=30
q TACHYON_B200_ARENA=0
q TACHYON_B200_ARENA=0 TACHYON_B200_PAD_MB=62
cat $out/${tag}_arena.log
