#!/bin/bash
# Refresh of the BN254 part of profiles/r2_y_* after the inlined running-sum kernel:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_run_r2c.sh r2yy'
tag=${1:-r2yy}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --workload groth16 --no-cpu-baseline > $out/${tag}_groth1.json 2> $out/${tag}_groth1.err
python bench.py --workload commit_batch --precompute > $out/${tag}_commit_pre.json 2> $out/${tag}_commit_pre.err
python bench.py --workload commit_batch > $out/${tag}_commit.json 2> $out/${tag}_commit.err
python tools/quick_gpu.py bn254 12,14,16,17,18,19,20,21,22,23,24 > $out/${tag}_sizes_bn254.log 2>&1
python tools/quick_gpu.py bn254 24 dist=witness > $out/${tag}_witness24.log 2>&1
export_rep() { ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2> /dev/null; rm -f $out/$1.ncu-rep; }
cmd="python bench.py --steps 2 --warmup 3 --no-parity --no-cpu-baseline --no-extra"
$cmd > $out/${tag}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
      --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu1.log 2>&1
$cmd > $out/${tag}_plain2.log 2>&1 &&
  ncu --set full --clock-control none \
      -k regex:'accumulate_kernel|reduce_blocks_kernel|reduce_tree_kernel|window_combine_kernel|fine_scatter_kernel|coarse_scatter_kernel|digits_coarse_hist_kernel|fine_hist_kernel|scan_apply_build_tasks_kernel' \
      --launch-skip 33 --launch-count 11 -o $out/${tag}_top -f $cmd > $out/${tag}_ncu2.log 2>&1
export_rep ${tag}_top
ls -la $out/${tag}_* | awk '{print $5, $9}'
