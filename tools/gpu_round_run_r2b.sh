#!/bin/bash
# Final round-2 measurement set behind profiles/r2_y_*: ONE B200 through gpurun, e.g.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round_run_r2b.sh r2y'
# 1. parity tests, 2. plain bench lines (never under a profiler), 3. ncu launch list of the default
# bench command, 4. ncu --set full of the heaviest kernels.  Outputs: gpurun_out/<tag>_*.
tag=${1:-r2y}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/${tag}_pytest.log

python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_ref.json 2> $out/${tag}_ref.err
python bench.py --curve bls12_381 --log-n 22 --no-cpu-baseline --no-extra > $out/${tag}_bls22.json 2> $out/${tag}_bls22.err
python bench.py --curve bn254_g2 --log-n 20 --no-cpu-baseline --no-extra > $out/${tag}_g2_20.json 2> $out/${tag}_g2_20.err
python bench.py --workload groth16 --no-cpu-baseline > $out/${tag}_groth1.json 2> $out/${tag}_groth1.err
python bench.py --workload commit_batch --precompute > $out/${tag}_commit_pre.json 2> $out/${tag}_commit_pre.err
python tools/quick_gpu.py bn254 12,14,16,17,18,19,20,21,22,23,24 > $out/${tag}_sizes_bn254.log 2>&1
TACHYON_B200_ARENA=0 python tools/quick_gpu.py bn254 21,24 > $out/${tag}_sizes_bn254_noarena.log 2>&1
python tools/quick_gpu.py bls12_381 16,18,19,20,22 > $out/${tag}_sizes_bls12_381.log 2>&1
python tools/quick_gpu.py bn254_g2 16,18,20 > $out/${tag}_sizes_bn254_g2.log 2>&1
python tools/quick_gpu.py bls12_381_g2 16,18,20 > $out/${tag}_sizes_bls12_381_g2.log 2>&1
timeout 300 python tools/fuzz_gpu.py 500 7 --g2 > $out/${tag}_fuzz.log 2>&1

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
cmd3="python tools/quick_gpu.py bls12_381 22"
$cmd3 > $out/${tag}_plain3.log 2>&1 &&
  ncu --set full --clock-control none -k regex:'accumulate_lockstep_kernel|reduce_blocks_kernel' \
      --launch-skip 6 --launch-count 2 -o $out/${tag}_bls_acc -f $cmd3 > $out/${tag}_ncu3.log 2>&1
export_rep ${tag}_bls_acc
cmd4="python tools/quick_gpu.py bn254_g2 20"
$cmd4 > $out/${tag}_plain4.log 2>&1 &&
  ncu --set full --clock-control none -k regex:'accumulate_pair_kernel|reduce_blocks_pair_kernel' \
      --launch-skip 6 --launch-count 2 -o $out/${tag}_g2_acc -f $cmd4 > $out/${tag}_ncu4.log 2>&1
export_rep ${tag}_g2_acc
cmd5="python tools/quick_gpu.py bls12_381_g2 20"
$cmd5 > $out/${tag}_plain5.log 2>&1 &&
  ncu --set full --clock-control none -k regex:'accumulate_pair_kernel|reduce_blocks_pair_kernel' \
      --launch-skip 6 --launch-count 2 -o $out/${tag}_blsg2_acc -f $cmd5 > $out/${tag}_ncu5.log 2>&1
export_rep ${tag}_blsg2_acc
ls -la $out/${tag}_* | awk '{print $5, $9}'
