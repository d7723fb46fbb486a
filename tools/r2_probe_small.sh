#!/bin/bash
# launch lists of small / shard-sized MSMs (where fixed costs dominate)
out=gpurun_out
python tools/quick_gpu.py bn254 16,18,20,21,24 > $out/r2a_quick.log 2>&1
for lg in 16 20 21; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/r2a_launches_$lg.csv python tools/quick_gpu.py bn254 $lg > $out/r2a_ncu_$lg.log 2>&1
done
tail -8 $out/r2a_quick.log
