#!/bin/bash
out=gpurun_out
tag=${1:-r2l}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
timeout 600 python tools/fuzz_gpu.py 1500 11 > $out/${tag}_fuzz.log 2>&1
timeout 400 python tools/fuzz_gpu.py 400 12 --g2 > $out/${tag}_fuzz_g2.log 2>&1
tail -3 $out/${tag}_pytest.log; tail -5 $out/${tag}_fuzz.log; tail -5 $out/${tag}_fuzz_g2.log
