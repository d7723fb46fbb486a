#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2u}
for pad in 0 2 6 14 30 62 126; do
  echo "== pad $pad MB" >> $out/${tag}_pad.log
  TACHYON_B200_PAD_MB=$pad timeout 300 python tools/quick_gpu.py bn254 24 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_pad.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python tools/quick_gpu.py bn254 24 > $out/${tag}_ncu.log 2>&1
cat $out/${tag}_pad.log
