#!/bin/bash
out=gpurun_out
tag=${1:-r2c}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python tools/quick_gpu.py bn254 16,18,20,21,24 > $out/${tag}_quick.log 2>&1
python tools/quick_gpu.py bn254 16,20,21,24 low_windows=0 > $out/${tag}_quick_nosplit.log 2>&1
python tools/quick_gpu.py bls12_381 19,22 > $out/${tag}_quick_bls.log 2>&1
python tools/quick_gpu.py bn254_g2 20 > $out/${tag}_quick_g2.log 2>&1
for lg in 21; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/${tag}_launches_$lg.csv python tools/quick_gpu.py bn254 $lg > $out/${tag}_ncu_$lg.log 2>&1
done
tail -5 $out/${tag}_pytest.log; cat $out/${tag}_quick.log $out/${tag}_quick_nosplit.log $out/${tag}_quick_bls.log $out/${tag}_quick_g2.log
