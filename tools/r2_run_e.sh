#!/bin/bash
out=gpurun_out
tag=${1:-r2e}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python tools/quick_gpu.py bn254 16,18,20,21,24 > $out/${tag}_quick.log 2>&1
python tools/quick_gpu.py bls12_381 19,22 > $out/${tag}_quick_bls.log 2>&1
python tools/quick_gpu.py bn254_g2 20 > $out/${tag}_quick_g2.log 2>&1
TACHYON_B200_TRACE=1 python tools/quick_gpu.py bn254 21,24 2>&1 | grep -E "window groups" | tail -4 > $out/${tag}_trace.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_ref.json 2> $out/${tag}_ref.err
tail -5 $out/${tag}_pytest.log; grep -h "^2\^" $out/${tag}_quick.log $out/${tag}_quick_bls.log $out/${tag}_quick_g2.log; cat $out/${tag}_trace.log; tail -c 600 $out/${tag}_bench.err; head -c 3000 $out/${tag}_bench.json; echo; cat $out/${tag}_ref.json | head -c 1500
