#!/bin/bash
# usage: tools/bench_env.sh "name ENV1=a ENV2=b" "name2 ..." ; runs a short bench per setting, output in gpurun_out/
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; v=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/b_$v.json 2> gpurun_out/b_$v.err
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
