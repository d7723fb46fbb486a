#!/bin/bash
out=gpurun_out
tag=${1:-r2j}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python tools/quick_gpu.py bn254 12,14,15,16,17,18,19,20,21,22,23,24 > $out/${tag}_quick.log 2>&1
python tools/quick_gpu.py bls12_381 16,18,19,20,22 > $out/${tag}_quick_bls.log 2>&1
python tools/quick_gpu.py bn254_g2 16,18,20 > $out/${tag}_quick_g2.log 2>&1
python bench.py --workload groth16 --steps 5 > $out/${tag}_groth16.json 2> $out/${tag}_groth16.err
python bench.py --workload commit_batch --steps 5 --log-n 16 --batch 64 > $out/${tag}_commit16.json 2> $out/${tag}_commit16.err
python bench.py --workload commit_batch --steps 5 --log-n 18 --batch 16 > $out/${tag}_commit18.json 2> $out/${tag}_commit18.err
tail -5 $out/${tag}_pytest.log; grep -h "^2\^" $out/${tag}_quick.log $out/${tag}_quick_bls.log $out/${tag}_quick_g2.log | cut -c1-170
for f in groth16 commit16 commit18; do python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$f.json"))
    print("$f", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), d["parity"], d.get("proof"))
except Exception as e:
    print("$f", "failed", e); print(open("$out/${tag}_$f.err").read()[-800:])
PY
done
