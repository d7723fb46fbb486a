#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-r2r}
timeout 600 python -m pytest tests -m gpu -x -q -k "lockstep or code_shapes or variants" 2>&1 | tail -5 > $out/${tag}_pytest.log
q() { echo "== $*" >> $out/${tag}_lock.log; timeout 300 python tools/quick_gpu.py "$@" 2>&1 | grep -v "^imad\|^generated" >> $out/${tag}_lock.log; }
q bls12_381 19,22
q bls12_381 19,22 acc_variant=3
q bn254 21,24 acc_variant=3
q bls12_381 22 reduce_roll=0
export_rep() { ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2> /dev/null; rm -f $out/$1.ncu-rep; }
cmd4="python tools/quick_gpu.py bn254_g2 20"
ncu --set full --clock-control none -k regex:'accumulate_pair_kernel|reduce_blocks_kernel' \
      --launch-skip 6 --launch-count 2 -o $out/${tag}_g2_acc -f $cmd4 > $out/${tag}_ncu4.log 2>&1
export_rep ${tag}_g2_acc
cmd3="python tools/quick_gpu.py bls12_381 22 acc_variant=3"
ncu --set full --clock-control none -k regex:'accumulate_lockstep_kernel' \
      --launch-skip 3 --launch-count 1 -o $out/${tag}_bls_lock -f $cmd3 > $out/${tag}_ncu3.log 2>&1
export_rep ${tag}_bls_lock
cat $out/${tag}_pytest.log $out/${tag}_lock.log
