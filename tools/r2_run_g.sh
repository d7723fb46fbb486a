#!/bin/bash
out=gpurun_out
tag=${1:-r2g}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python tools/quick_gpu.py bn254 20,21,24 stage_points=1 > $out/${tag}_quick_staged.log 2>&1
python tools/quick_gpu.py bls12_381 19,22 stage_points=1 > $out/${tag}_quick_bls_staged.log 2>&1
python tools/quick_gpu.py bn254_g2 20 stage_points=1 > $out/${tag}_quick_g2_staged.log 2>&1
python tools/quick_gpu.py bls12_381_g2 18 stage_points=1 > $out/${tag}_quick_g2b_staged.log 2>&1
python tools/quick_gpu.py bls12_381_g2 18 > $out/${tag}_quick_g2b.log 2>&1
python bench.py --workload commit_batch --steps 5 > $out/${tag}_commit.json 2> $out/${tag}_commit.err
python bench.py --workload commit_batch --steps 5 --precompute > $out/${tag}_commit_pre.json 2> $out/${tag}_commit_pre.err
python bench.py --workload commit_batch --steps 5 --log-n 16 --batch 64 > $out/${tag}_commit16.json 2> $out/${tag}_commit16.err
python bench.py --workload commit_batch --steps 5 --log-n 16 --batch 64 --precompute > $out/${tag}_commit16_pre.json 2> $out/${tag}_commit16_pre.err
tail -5 $out/${tag}_pytest.log; grep -h "^2\^" $out/${tag}_quick_staged.log $out/${tag}_quick_bls_staged.log $out/${tag}_quick_g2_staged.log $out/${tag}_quick_g2b_staged.log $out/${tag}_quick_g2b.log
for f in commit commit_pre commit16 commit16_pre; do python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$f.json"))
    print("$f", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), d["parity"], d["config"].get("register_ms"), d["config"].get("window_bits"), d["config"].get("windows"))
except Exception as e:
    print("$f", "failed", e); print(open("$out/${tag}_$f.err").read()[-800:])
PY
done
