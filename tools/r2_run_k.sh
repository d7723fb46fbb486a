#!/bin/bash
out=gpurun_out
tag=${1:-r2k}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python tools/quick_gpu.py bn254 12,14,16,18,20,22 > $out/${tag}_quick.log 2>&1
timeout 300 python tools/fuzz_gpu.py 400 7 > $out/${tag}_fuzz.log 2>&1
timeout 300 python tools/fuzz_gpu.py 150 8 --g2 > $out/${tag}_fuzz_g2.log 2>&1
timeout 120 compute-sanitizer --tool racecheck --kernel-regex kns=reduce_blocks python tools/quick_gpu.py bn254 12 > $out/${tag}_racecheck.log 2>&1
tail -3 $out/${tag}_pytest.log; grep -h "^2\^" $out/${tag}_quick.log | cut -c1-150; tail -3 $out/${tag}_fuzz.log; tail -3 $out/${tag}_fuzz_g2.log; tail -6 $out/${tag}_racecheck.log
