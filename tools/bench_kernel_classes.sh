#!/bin/bash
# short bench per library variant, printing only the chosen classes: tools/bench_kernel_classes.sh "rk_mom fft" default v1 v2 ...
mkdir -p gpurun_out
pat=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then L=""; else L=$PWD/snac_b200/variants/libsnacb200_$v.so; fi
  SNAC_B200_LIB=$L python bench.py --steps 12 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/b_$v.json 2> gpurun_out/b_$v.err
  python - "$v" "$pat" <<'PY'
import json, sys
v, pat = sys.argv[1], sys.argv[2].split()
try:
    d = json.load(open(f"gpurun_out/b_{v}.json")); c = d["roofline"]["classes"]
    print(v, d["ms_per_step"], {k: (x["ms_per_step"], x.get("achieved_gbs")) for k, x in c.items() if any(p in k for p in pat)})
except Exception as e:
    print(v, "FAILED", e); print(open(f"gpurun_out/b_{v}.err").read()[-800:])
PY
done
