#!/bin/bash
# two GPUs: the NCCL path of bench.py (join_ranks), default and device ladder
out=gpurun_out
tag=${1:-r2h}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $out/${tag}_bench2.json 2> $out/${tag}_bench2.err
echo "rc=$?"; tail -c 1500 $out/${tag}_bench2.err
python bench.py --workload commit_batch --steps 5 --precompute > $out/${tag}_commit_pre.json 2> $out/${tag}_commit_pre.err
python bench.py --workload commit_batch --steps 5 > $out/${tag}_commit.json 2> $out/${tag}_commit.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench2.json"))
print("N=2 ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["parity"], d["stages_ms"])
for k in d:
    if k.startswith('extra'):
        e=d[k]
        print(k, {x:(round(v,3) if isinstance(v,float) else v) for x,v in e.items() if x in ('ms_per_msm','ms_per_step','e2e_ms_per_msm','e2e_ms_per_step','imad_frac','parity','devices')})
for f in ("commit","commit_pre"):
    d=json.load(open("$out/${tag}_%s.json"%f)); print(f, round(d["ms_per_step"],2), d["config"]["stages_ms_per_commitment"])
PY
