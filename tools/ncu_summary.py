"""Summarise ncu output for profiles/:
  python tools/ncu_summary.py launches <launch_list.csv>            -> per-kernel table (count, total us, share, DRAM bytes)
  python tools/ncu_summary.py full <report.ncu-rep> [out.json]      -> selected `--set full` metrics per captured launch
"""
import collections
import csv
import json
import re
import subprocess
import sys


def _num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def launches(path, out_json=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        kid = row["ID"]
        e = per.setdefault(kid, {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")})
        v = _num(row["Metric Value"])
        unit = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        elif row["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            e[row["Metric Name"]] = v * mult
    agg = {}
    for e in per.values():
        a = agg.setdefault(e["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += e.get("us", 0.0)
        a[2] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"{len(per)} launches, {tot / 1e3:.2f} ms serialised")
    print("| kernel | launches | total us | share | DRAM MB / launch |")
    print("|---|---|---|---|---|")
    for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {n} | {us:.1f} | {100 * us / tot:.1f} % | {by / n / 1e6:.1f} |")
    if out_json:
        gem = {k: v for k, v in agg.items() if "gemm" in k or "conv64_halo" in k}   # the tcgen05 GEMM family of bench.py's roofline
        n = sum(v[0] for v in gem.values())
        doc = {"source": path, "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, one eager train step "
                                       "(B=128, L=128), ncu --clock-control none",
               "gemm_launches": n, "gemm_dram_bytes_per_launch": sum(v[2] for v in gem.values()) / max(n, 1),
               "gemm_time_share": sum(v[1] for v in gem.values()) / tot,
               "per_kernel": {k: {"launches": v[0], "total_us": round(v[1], 1), "dram_bytes_per_launch": v[2] / v[0]}
                              for k, v in agg.items()}}
        open(out_json, "w").write(json.dumps(doc, indent=1) + "\n")


WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def full(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in data:
        e = {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")}
        for w in WANT:
            if w in idx:
                e[w] = f"{r[idx[w]]} {units[idx[w]]}".strip()
        res.append(e)
    txt = json.dumps(res, indent=1)
    if out:
        open(out, "w").write(txt + "\n")
    else:
        print(txt)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
