#!/bin/bash
# k concurrent bare H2D probes, one per GPU, for k = 1, 2, 4, 8 (as many as the box has), in the
# three host-memory modes of tools/probe/h2d_probe.cu; with numactl when available, each process is
# also tried bound to the NUMA node nvidia-smi reports for its GPU.  Output: gpurun_out/<tag>_h2d.txt
tag=${1:-r2}
out=gpurun_out/${tag}_h2d.txt
nvcc -O2 -o tools/probe/h2d_probe tools/probe/h2d_probe.cu || exit 1
n=$(nvidia-smi -L | wc -l)
{
  nvidia-smi topo -m
  echo "numactl: $(command -v numactl || echo absent)"; lscpu | grep -i -E "numa|model name|socket"
  for mode in 0 1 2; do
    for k in 1 2 4 8; do
      [ $k -gt $n ] && continue
      echo "== $k concurrent processes, mode $mode"
      start=$(python3 -c "import time; print(time.time() + 3)")
      for ((g = 0; g < k; g++)); do tools/probe/h2d_probe $g 256 20 $mode $start & done
      wait
    done
  done
  if command -v numactl > /dev/null; then
    for k in 4 8; do
      [ $k -gt $n ] && continue
      echo "== $k concurrent processes, mode 2, numactl --cpunodebind/--membind to the GPU's node"
      start=$(python3 -c "import time; print(time.time() + 3)")
      for ((g = 0; g < k; g++)); do
        node=$(nvidia-smi topo -m | awk -v g="GPU$g" '$1 == g {print $(NF-1)}')
        numactl --cpunodebind=${node:-0} --membind=${node:-0} tools/probe/h2d_probe $g 256 20 2 $start &
      done
      wait
    done
  fi
} > $out 2>&1
cat $out
