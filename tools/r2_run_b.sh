#!/bin/bash
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/r2b_pytest.log
python tools/quick_gpu.py bn254 16,18,20,21,24 > $out/r2b_quick.log 2>&1
python tools/quick_gpu.py bn254 16,20,21,24 low_windows=0 > $out/r2b_quick_nosplit.log 2>&1
python tools/quick_gpu.py bls12_381 19,22 > $out/r2b_quick_bls.log 2>&1
python tools/quick_gpu.py bn254_g2 20 > $out/r2b_quick_g2.log 2>&1
tail -5 $out/r2b_pytest.log; cat $out/r2b_quick.log $out/r2b_quick_nosplit.log $out/r2b_quick_bls.log $out/r2b_quick_g2.log
