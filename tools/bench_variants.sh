#!/bin/bash
# usage: tools/bench_variants.sh name1 name2 ...   ("default" = the in-tree library); output in gpurun_out/
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 1"
for v in "$@"; do
  if [ "$v" = default ]; then $B > gpurun_out/b_$v.json 2> gpurun_out/b_$v.err
  else SNAC_B200_LIB=$PWD/snac_b200/variants/libsnacb200_$v.so $B > gpurun_out/b_$v.json 2> gpurun_out/b_$v.err; fi
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/b_{v}.json")); c = d["roofline"]["classes"]
    print(v, d["ms_per_step"], d["config"]["mg_vcycles_per_solve"],
          {k: (x["ms_per_step"], x.get("achieved_gbs")) for k, x in c.items() if x["ms_per_step"] > 0})
except Exception as e:
    print(v, "FAILED", e); print(open(f"gpurun_out/b_{v}.err").read()[-1500:])
PY
done
