#!/bin/bash
out=gpurun_out
tag=${1:-r2i}
for wb in 12 13 14 15 16; do echo "wb=$wb"; python tools/quick_gpu.py bn254 14,16,17 window_bits=$wb | grep "^2\^"; done > $out/${tag}_small_wb.log 2>&1
for wb in 14 15 16 17; do echo "wb=$wb"; python tools/quick_gpu.py bn254 18,19 window_bits=$wb | grep "^2\^"; done >> $out/${tag}_small_wb.log 2>&1
for wb in 16 17 18; do echo "wb=$wb"; python tools/quick_gpu.py bn254 20 window_bits=$wb | grep "^2\^"; done >> $out/${tag}_small_wb.log 2>&1
cat $out/${tag}_small_wb.log
