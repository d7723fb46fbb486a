#!/bin/bash
export TACHYON_B200_TRACE=1
for lg in 21 24 20 16; do
python tools/quick_gpu.py bn254 $lg 2>&1 | grep -E "window groups|^2\^" | tail -3
done
python tools/quick_gpu.py bls12_381 22 2>&1 | grep -E "window groups|^2\^" | tail -3
