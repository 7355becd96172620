#!/bin/bash
# usage: tools/sweep.sh "<bench args>" name1 name2 ... ; runs a short bench per library variant built by tools/build_variant.py
# (snac_b200/variants/libsnacb200_<name>.so; "base" = the default library) and prints the per-class times
args="$1"; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset SNAC_B200_LIB; else export SNAC_B200_LIB=$PWD/snac_b200/variants/libsnacb200_$v.so; fi
  python bench.py $args --no-cpu --no-alt --no-check --e2e-steps 1 > gpurun_out/s_$v.json 2> gpurun_out/s_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/s_{v}.json").read().strip().splitlines()[-1]); c = d["roofline"]["classes"]
    print(v, d["ms_per_step"], {k: (x["ms_per_step"], x.get("achieved_gbs")) for k, x in c.items() if x["ms_per_step"] > 0})
except Exception as e:
    print(v, "FAILED", e); print(open(f"gpurun_out/s_{v}.err").read()[-800:])
PY
done
