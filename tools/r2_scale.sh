#!/bin/bash
# 8-GPU box: the scaling line at N = 8 and N = 4 (+ the extras bench.py carries at N > 1) and the
# bare H2D probe.  gpurun --gpus 8 -- 'bash tools/r2_scale.sh r2s'
out=gpurun_out
tag=${1:-r2s}
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench$n.json 2> $out/${tag}_bench$n.err
  echo "N=$n rc=$?"
done
[ -z "$SKIP_PROBE" ] && bash tools/h2d_probe.sh $tag > /dev/null 2>&1
python - <<PY
import json
for n in (8, 4):
    try:
        d=json.load(open("$out/${tag}_bench%d.json" % n))
    except Exception as e:
        print(n, "failed", e); print(open("$out/${tag}_bench%d.err" % n).read()[-1500:]); continue
    print("N=%d ms %.3f e2e %.3f h2d GB/s/gpu %s parity %s stages %s" % (n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("h2d_gbs_per_gpu"), d["parity"], {k: round(v,3) for k,v in d["stages_ms"].items()}))
    for k in d:
        if k.startswith('extra'):
            e=d[k]
            print("  ", k, {x:(round(v,3) if isinstance(v,float) else v) for x,v in e.items() if x in ('ms_per_msm','ms_per_step','e2e_ms_per_msm','e2e_ms_per_step','imad_frac','parity','devices')})
PY
[ -z "$SKIP_PROBE" ] && grep -E "==|device" $out/${tag}_h2d.txt | head -80
