"""Turn the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (see profiles/README.md).
   python tools/summarize_profiles.py [round-tag]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")


def launches():
    src = os.path.join(ROOT, "gpurun_out", f"{TAG}_unet_launches.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    i_name, i_metric, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    for r in rows[1:]:
        k = per.setdefault(r[i_id], {"kernel": r[i_name]})
        k[r[i_metric]] = float(r[i_val].replace(",", ""))
    agg = {}
    for k in per.values():
        name = re.sub(r"\(.*", "", k["kernel"]).replace("dm::", "").strip()
        name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, {"kernel": name, "launches": 0, "us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0})
        a["launches"] += 1
        a["us"] += k.get("gpu__time_duration.sum", 0.0) / 1e3
        a["dram_read_MB"] += k.get("dram__bytes_read.sum", 0.0) / 1e6
        a["dram_write_MB"] += k.get("dram__bytes_write.sum", 0.0) / 1e6
    tot = sum(a["us"] for a in agg.values())
    ks = sorted(agg.values(), key=lambda a: -a["us"])
    for a in ks:
        a["share"] = round(a["us"] / tot, 4)
        for f in ("us", "dram_read_MB", "dram_write_MB"):
            a[f] = round(a[f], 1)
    bf = os.environ.get("DM_BF", "54")
    json.dump({"what": f"ncu (--clock-control none), one {bf}-forward U-Net micro-batch @64x64, eager replay; per-launch times are "
                       "cold-cache and serialised -> compare shares", "total_us": round(tot, 1), "kernels": ks},
              open(os.path.join(OUT, f"{TAG}_unet_launch_summary.json"), "w"), indent=1)
    ig = [a for a in ks if a["kernel"].startswith("igemm_kernel")]
    n = sum(a["launches"] for a in ig)
    rd, wr = sum(a["dram_read_MB"] for a in ig) * 1e6, sum(a["dram_write_MB"] for a in ig) * 1e6
    json.dump({"dram_bytes_per_launch": (rd + wr) / n, "launches": n, "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
               "micro_batch_forwards": int(bf),
               "source": f"profiles/{TAG}_unet_launches.csv: ncu dram__bytes_read.sum + dram__bytes_write.sum over the {n} igemm_kernel "
                         f"launches of one {bf}-forward U-Net micro-batch @64x64 (cold cache per launch)"},
              open(os.path.join(OUT, f"{TAG}_igemm_traffic.json"), "w"), indent=1)
    # keep the raw list too (trimmed to the three metrics)
    with open(os.path.join(OUT, f"{TAG}_unet_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum [ns]", "dram__bytes_read.sum [B]", "dram__bytes_write.sum [B]"])
        i_grid, i_block = hdr.index("Grid Size"), hdr.index("Block Size")
        seen = {}
        for r in rows[1:]:
            seen.setdefault(r[i_id], [r[i_id], re.sub(r"\(dm::.*", "", r[i_name]), r[i_grid], r[i_block], None, None, None])
            j = {"gpu__time_duration.sum": 4, "dram__bytes_read.sum": 5, "dram__bytes_write.sum": 6}.get(r[i_metric])
            if j:
                seen[r[i_id]][j] = r[i_val].replace(",", "")
        for v in seen.values():
            w.writerow(v)
    print("launch summary:", [(a["kernel"], a["launches"], a["us"], a["share"]) for a in ks[:8]])


def full():
    rep = os.path.join(ROOT, "gpurun_out", f"{TAG}_ops_full.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__cluster_size" if "launch__cluster_size" in hdr else "launch__grid_size",
            "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]
    targets = ["conv3x3 320->320 @64x64 x32", "GEGLU 320->2560 @64x64 x32", "1x1 320->320 +res @64x64 x32", "conv3x3 1280->1280 @16x16 x32",
               "conv3x3 1280->1280 @32x32 x32 (CTA pairs)", "self-attn d=40 T=4096 x32", "self-attn d=80 T=1024 x32",
               "cross-attn d=40 T=4096 x 77 keys x32", "GroupNorm+SiLU 320ch @64x64 x32", "LayerNorm 320 x 131072 rows"]
    ks = []
    i_name = hdr.index("Kernel Name")
    for n, r in enumerate(rows[2:]):
        e = {"target": targets[n] if n < len(targets) else f"launch {n}", "kernel": re.sub(r"\(dm::.*|\(.*", "", r[i_name])}
        for m in want:
            if m in hdr:
                j = hdr.index(m)
                try:
                    e[f"{m} [{units[j]}]" if units[j] else m] = float(r[j].replace(",", ""))
                except ValueError:
                    e[m] = r[j]
        ks.append(e)
    json.dump({"what": "ncu --set full --clock-control none, kernels launched alone through the op-level ABI after a 256 MB L2 flush "
                       f"(tools/profile_target.py ops); report gpurun_out/{TAG}_ops_full.ncu-rep (not committed: 20+ MB)", "kernels": ks},
              open(os.path.join(OUT, f"{TAG}_ops_full_summary.json"), "w"), indent=1)
    for e in ks:
        print(e["target"], "|", e["kernel"], "|", {k.split(".")[0] + "." + k.split(" ")[-1]: v for k, v in e.items() if "duration" in k or "tensor" in k or "dram_thr" in k or "pipe_xu" in k})


if __name__ == "__main__":
    launches()
    full()
