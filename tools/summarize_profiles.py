"""Summaries of the ncu captures for profiles/: python tools/summarize_profiles.py [round-tag]
 - gpurun_out/launches.csv   (ncu --metrics gpu__time_duration.sum)  -> profiles/<tag>_ncu_launches.csv + _summary.md
 - gpurun_out/r01_full.ncu-rep (ncu --set full)                      -> profiles/<tag>_ncu_full_summary.md, ncu_traffic.json
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")


def short(name):
    name = name.replace("void snac::", "").replace("snac::", "")
    return re.sub(r"\(.*", "", name)


# ---- launch list
src = os.path.join(ROOT, "gpurun_out", "launches.csv")
rows = list(csv.reader(open(src)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
idx = {k: j for j, k in enumerate(rows[h])}
tot = collections.OrderedDict()
for r in rows[h + 1:]:
    try:
        v = float(r[idx["Metric Value"]]) / 1000.0
    except (ValueError, IndexError):
        continue
    k = short(r[idx["Kernel Name"]])
    n, t = tot.get(k, (0, 0.0))
    tot[k] = (n + 1, t + v)
shutil.copy(src, os.path.join(out, f"{tag}_ncu_launches.csv"))
total = sum(t for _, t in tot.values())
with open(os.path.join(out, f"{tag}_ncu_launches_summary.md"), "w") as f:
    f.write(f"# ncu launch list of `bench.py --steps 2 --warmup 4` (TGV 256^3, 1 GPU), {sum(n for n, _ in tot.values())} launches\n\n")
    f.write("Command (profiling recipe): `ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 400 --csv "
            "--log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 4 --no-cpu --e2e-steps 1`\n\n")
    f.write("Per-launch times are cold-cache and serialised: compare shares, not absolutes.\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {n} | {t:.1f} | {t / total:.3f} |\n")

# ---- full capture
rep = os.path.join(ROOT, "gpurun_out", "r01_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
ix = {k: j for j, k in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram read MB"), ("dram__bytes_write.sum", "dram write MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp instr"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
cols = [(c, t) for c, t in cols if c in ix]
traffic = {}
with open(os.path.join(out, f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# ncu --set full capture ({tag}), TGV 256^3, 1 GPU\n\n")
    f.write("`ncu --set full --clock-control none --import-source on -k regex:\"mg_ring|mg_coarse|rk_mom|r2r_radix|correc|fillps\" "
            "-s 60 -c 20 python bench.py --steps 1 --warmup 4 --no-cpu --e2e-steps 1`\n\n")
    f.write("| kernel | grid | " + " | ".join(t for _, t in cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for r in data:
        name = short(r[ix["Kernel Name"]])
        vals = []
        for c, _ in cols:
            try:
                vals.append(f"{float(r[ix[c]]):.1f}")
            except ValueError:
                vals.append(r[ix[c]])
        f.write(f"| {name} | {r[ix['Grid Size']]} | " + " | ".join(vals) + " |\n")
        try:
            by = (float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])) * 1e6
        except ValueError:
            continue
        key = "mg_down_l0" if "mg_ring_down" in name else ("mg_up_l0" if "mg_ring_up" in name else None)
        if key and by > traffic.get(key, 0.0):
            traffic[key] = by   # the full-size launch (all planes iterating)
json.dump(traffic, open(os.path.join(out, "ncu_traffic.json"), "w"))
print(open(os.path.join(out, f"{tag}_ncu_launches_summary.md")).read())
print(open(os.path.join(out, f"{tag}_ncu_full_summary.md")).read())
print(traffic)
