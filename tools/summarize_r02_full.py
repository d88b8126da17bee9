"""gpurun_out/r02_ops_full.ncu-rep / r02_attn3_d40.ncu-rep (ncu --set full) -> profiles/r02_ops_full_summary.json.
   python tools/summarize_r02_full.py"""
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__cluster_size", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]
TARGETS = ["conv3x3 320->320 @64x64 x32", "GEGLU 320->2560 @64x64 x32", "1x1 320->320 +res @64x64 x32", "conv3x3 1280->1280 @16x16 x32",
           "conv3x3 1280->1280 @32x32 x32 (CTA pairs)", "self-attn d=40 T=4096 x32", "self-attn d=80 T=1024 x32",
           "cross-attn d=40 T=4096 x 77 keys x32", "GroupNorm+SiLU 320ch @64x64 x32 (stand-alone)", "LayerNorm 320 x 131072 rows",
           "conv3x3 320->320 @64x64 x32 + GroupNorm statistics in the epilogue", "GroupNorm+SiLU 320ch @64x64 x32 (fold + apply)"]


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return rows[0], rows[1], rows[2:]


def entry(hdr, units, r, target):
    e = {"target": target, "kernel": re.sub(r"\(dm::.*|\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "")}
    for m in WANT:
        if m in hdr:
            j = hdr.index(m)
            try:
                e[f"{m} [{units[j]}]" if units[j] else m] = float(r[j].replace(",", ""))
            except ValueError:
                e[m] = r[j]
    return e


def main():
    ks = []
    rep = os.path.join(ROOT, "gpurun_out", "r02_ops_full.ncu-rep")
    if os.path.exists(rep):
        hdr, units, rows = rows_of(rep)
        ks += [entry(hdr, units, r, TARGETS[n] if n < len(TARGETS) else f"launch {n}") for n, r in enumerate(rows)]
    rep = os.path.join(ROOT, "gpurun_out", "r02_attn3_d40.ncu-rep")
    if os.path.exists(rep):
        hdr, units, rows = rows_of(rep)
        ks += [entry(hdr, units, r, "self-attn d=40 T=4096 x8 (profiles/r02_attn3_d40.ncu-rep, source-level stalls inside)") for r in rows]
    json.dump({"what": "ncu --set full --clock-control none, kernels launched alone through the op-level ABI after a 256 MB L2 flush "
                       "(tools/profile_target.py ops / attn)", "kernels": ks}, open(os.path.join(OUT, "r02_ops_full_summary.json"), "w"), indent=1)
    for e in ks:
        print(e["target"], "|", e["kernel"], "|", {k.split(".")[0][:24]: round(v, 3) for k, v in e.items() if isinstance(v, float) and
                                                   ("duration" in k or "tensor_cycles" in k or "dram_thr" in k or "pipe_xu" in k or "issue_active" in k)})


if __name__ == "__main__":
    main()
