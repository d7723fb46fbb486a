#!/bin/bash
# Last check of the build that ships: parity suite, smoke, stage times of all four groups, bench line
out=gpurun_out; mkdir -p $out; tag=${1:-r2fin}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
q() { echo "== $*" >> $out/${tag}_q.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep "^2\^" | cut -c1-118 >> $out/${tag}_q.log; }
q bn254 16,20,21,23,24
q bls12_381 19,22
q bn254_g2 20
q bls12_381_g2 20
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cat $out/${tag}_pytest.log $out/${tag}_q.log; tail -1 $out/${tag}_smoke.log
