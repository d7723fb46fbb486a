#!/bin/bash
out=gpurun_out
tag=${1:-r2f}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python tools/quick_gpu.py bn254 16,18,20,21,24 > $out/${tag}_quick.log 2>&1
python tools/quick_gpu.py bn254 20,21,24 device_ladder=1 > $out/${tag}_quick_dev.log 2>&1
python tools/quick_gpu.py bls12_381 19,22 > $out/${tag}_quick_bls.log 2>&1
python tools/quick_gpu.py bn254_g2 20 > $out/${tag}_quick_g2.log 2>&1
python bench.py --steps 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -5 $out/${tag}_pytest.log; grep -h "^2\^" $out/${tag}_quick.log $out/${tag}_quick_dev.log $out/${tag}_quick_bls.log $out/${tag}_quick_g2.log; tail -c 600 $out/${tag}_bench.err
