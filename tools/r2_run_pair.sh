#!/bin/bash
# G2 accumulation variants: parity tests of the G2 groups, then stage timings per variant
out=gpurun_out; mkdir -p $out; tag=${1:-r2p}
timeout 600 python -m pytest tests -m gpu -x -q -k "g2" 2>&1 | tail -5 > $out/${tag}_pytest.log
for v in 0 1 2; do
  echo "== bn254_g2 acc_variant=$v" >> $out/${tag}_g2.log
  timeout 300 python tools/quick_gpu.py bn254_g2 16,18,20 acc_variant=$v 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_g2.log
done
for v in 0 1 2; do
  echo "== bls12_381_g2 acc_variant=$v" >> $out/${tag}_g2.log
  timeout 300 python tools/quick_gpu.py bls12_381_g2 18,20 acc_variant=$v 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_g2.log
done
cat $out/${tag}_pytest.log $out/${tag}_g2.log
