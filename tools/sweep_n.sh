#!/bin/bash
# usage: tools/sweep_n.sh N "<bench args>" name1 name2 ... ; like sweep.sh on N GPUs (torchrun), per-class times per library variant
n="$1"; args="$2"; shift; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset SNAC_B200_LIB; else export SNAC_B200_LIB=$PWD/snac_b200/variants/libsnacb200_$v.so; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n $args --no-cpu --no-alt --e2e-steps 1 > gpurun_out/sn${n}_$v.json 2> gpurun_out/sn${n}_$v.err
  python - "$v" "$n" <<'PY'
import json, sys
v, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/sn{n}_{v}.json").read().strip().splitlines()[-1]); c = d["roofline"]["classes"]
    print(v, d["ms_per_step"], d["check"]["rel_l2"] if d.get("check") else None, {k: (x["ms_per_step"], x.get("achieved_gbs")) for k, x in c.items() if x["ms_per_step"] > 0})
except Exception as e:
    print(v, "FAILED", e); print(open(f"gpurun_out/sn{n}_{v}.err").read()[-800:])
PY
done
