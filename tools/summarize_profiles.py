"""Summaries of the ncu captures for profiles/ (the captures themselves stay in gpurun_out/, which is scratch):

  python tools/summarize_profiles.py TAG --launches gpurun_out/X_launches.csv --full gpurun_out/X_prof.ncu-rep \
         --workload channel --cmd "the command that was profiled"

 - launch list (ncu --metrics gpu__time_duration.sum)  -> profiles/TAG_ncu_launches.csv + TAG_ncu_launches_summary.md
 - full capture (ncu --set full)                        -> profiles/TAG_ncu_full_summary.md, and the DRAM traffic per
   launch of each kernel class into profiles/ncu_traffic.json under "<workload>:<class>" (read by bench.py)
"""
import argparse
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

# kernel name -> the class name of snac_profile_name (bench.py's roofline classes)
CLASS_OF = [("rk_mom", "rk_mom"), ("fillps", "fillps"), ("correc", "correc_updt_pressure"), ("tri_solve", "direct_tridiag"),
            ("r2r_row", "direct_transform"), ("mg_ring_down", "mg_down_l0"), ("mg_ring_up", "mg_up_l0")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def short(name):
    name = name.replace("void snac::", "").replace("snac::", "").replace("void ", "")
    name = re.sub(r"\(anonymous namespace\)::|unnamed>::|<unnamed>::", "", name)
    return re.sub(r"\(.*", "", name)


def launches(tag, src, cmd):
    rows = list(csv.reader(open(src)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    idx = {k: j for j, k in enumerate(rows[h])}
    tot = collections.OrderedDict()
    for r in rows[h + 1:]:
        try:
            v = float(r[idx["Metric Value"]]) * UNIT.get(r[idx["Metric Unit"]], 1e-3)
        except (ValueError, IndexError, KeyError):
            continue
        k = short(r[idx["Kernel Name"]])
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + v)
    shutil.copy(src, os.path.join(OUT, f"{tag}_ncu_launches.csv"))
    total = sum(t for _, t in tot.values())
    with open(os.path.join(OUT, f"{tag}_ncu_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}), {sum(n for n, _ in tot.values())} launches\n\n")
        f.write(f"Command (profiling recipe): `{cmd}`\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare shares, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {t / total:.3f} |\n")


def full(tag, rep, cmd, workload, points):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: j for j, k in enumerate(hdr)}

    def val(r, c):
        try:
            return float(r[ix[c]]) * UNIT.get(units[ix[c]], 1.0)
        except (ValueError, KeyError):
            return float("nan")

    cols = [("gpu__time_duration.sum", "us", 1.0), ("dram__bytes_read.sum", "dram read MB", 1e-6), ("dram__bytes_write.sum", "dram write MB", 1e-6),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1.0),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", 1.0),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0), ("launch__registers_per_thread", "regs", 1.0),
            ("smsp__inst_executed.sum", "warp instr", 1.0), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %", 1.0),
            ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1.0), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0)]
    cols = [c for c in cols if c[0] in ix]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    traffic_path = os.path.join(OUT, "ncu_traffic.json")
    try:
        traffic = json.load(open(traffic_path))
    except (OSError, ValueError):
        traffic = {}
    seen = {}
    with open(os.path.join(OUT, f"{tag}_ncu_full_summary.md"), "w") as f:
        f.write(f"# ncu --set full capture ({tag})\n\n`{cmd}`\n\n")
        f.write("| kernel | grid | block | " + " | ".join(t for _, t, _ in cols) + " | top stalls (warps per issue) |\n")
        f.write("|---|---|---|" + "---|" * (len(cols) + 1) + "\n")
        for r in data:
            name = short(r[ix["Kernel Name"]])
            st = sorted(((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), val(r, h)) for h in stalls),
                        key=lambda kv: -kv[1])[:3]
            f.write(f"| {name} | {r[ix['Grid Size']]} | {r[ix['Block Size']]} | " + " | ".join(f"{val(r, c) * s:.1f}" for c, _, s in cols) +
                    " | " + ", ".join(f"{k} {v:.1f}" for k, v in st) + " |\n")
            for pat, cls in CLASS_OF:
                if pat in name:
                    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
                    if b > seen.get(cls, 0.0):     # the largest launch of the class
                        seen[cls] = b
        f.write("\nDRAM bytes per launch of the largest launch of each class (-> `ncu_traffic.json`): " +
                ", ".join(f"{k} {v / 1e6:.1f} MB" for k, v in seen.items()) + "\n")
    for k, v in seen.items():   # per grid point of the captured workload: bench.py scales it to the points of its launches
        traffic[f"{workload}:{k}"] = {"dram_bytes_per_point": v / points, "captured_on_points": points, "capture": f"{tag}_ncu_full_summary.md"}
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--full")
    ap.add_argument("--cmd", default="")
    ap.add_argument("--workload", default="channel")
    ap.add_argument("--points", type=float, default=1.0, help="grid points a full-field launch of the captured run covers")
    a = ap.parse_args()
    if a.launches:
        launches(a.tag, a.launches, a.cmd)
    if a.full:
        full(a.tag, a.full, a.cmd, a.workload, a.points)
