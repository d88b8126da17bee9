"""Round-2 ncu launch lists (gpurun_out/r02_<key>_launches.csv, one micro-batch of a benched plan) -> committed summaries:
   profiles/r02_<key>_launches.csv (trimmed raw list), profiles/r02_<key>_launch_summary.json (per-kernel totals / shares),
   profiles/r02_igemm_traffic.json[<config>] (DRAM bytes per igemm launch: bench.py's roofline.traffic).
   python tools/summarize_r02.py typ config2 54 2"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")


def main(key, config, Bf, aux):
    src = os.path.join(ROOT, "gpurun_out", f"r02_{key}_launches.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "Grid Size", "Block Size")}
    per = {}
    for r in rows[1:]:
        k = per.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        k[r[ix["Metric Name"]]] = v

    def short(n):
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*", "", n)
        return n.replace("dm::", "").strip()

    agg = {}
    for k in per.values():
        a = agg.setdefault(short(k["kernel"]), {"kernel": short(k["kernel"]), "launches": 0, "us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0})
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
    cls = {"igemm": 0.0, "attention": 0.0, "other": 0.0}
    for a in ks:
        c = "igemm" if a["kernel"].startswith("igemm_kernel") else "attention" if "attention" in a["kernel"] else "other"
        cls[c] += a["us"]
    json.dump({"what": f"ncu (--clock-control none, --profile-from-start off), ONE micro-batch of the benched {config} plan (Bf={Bf}, aux={aux}), eager "
                       "replay; per-launch times are cold-cache and serialised -> compare SHARES with bench.py's event-timed shares",
               "total_us": round(tot, 1), "class_shares": {k: round(v / tot, 4) for k, v in cls.items()}, "launches": len(per), "kernels": ks},
              open(os.path.join(OUT, f"r02_{key}_launch_summary.json"), "w"), indent=1)
    ig = [k for k in per.values() if short(k["kernel"]).startswith("igemm_kernel")]
    rd = sum(k.get("dram__bytes_read.sum", 0.0) for k in ig)
    wr = sum(k.get("dram__bytes_write.sum", 0.0) for k in ig)
    tp = os.path.join(OUT, "r02_igemm_traffic.json")
    d = json.load(open(tp)) if os.path.exists(tp) else {}
    d[config] = {"dram_bytes_per_launch": (rd + wr) / len(ig), "launches": len(ig), "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
                 "Bf": Bf, "aux": aux,
                 "source": f"profiles/r02_{key}_launches.csv: ncu dram__bytes_read.sum + dram__bytes_write.sum over the {len(ig)} igemm_kernel launches of "
                           f"one micro-batch of the benched plan (cold cache per launch)"}
    json.dump(d, open(tp, "w"), indent=1, sort_keys=True)
    with open(os.path.join(OUT, f"r02_{key}_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum [ns]", "dram__bytes_read.sum [B]", "dram__bytes_write.sum [B]"])
        for i in sorted(per):
            k = per[i]
            w.writerow([i, short(k["kernel"]) + re.sub(r"^[^<]*", "", re.sub(r"\(.*", "", k["kernel"]))[:0], k["grid"], k["block"],
                        int(k.get("gpu__time_duration.sum", 0)), int(k.get("dram__bytes_read.sum", 0)), int(k.get("dram__bytes_write.sum", 0))])
    print(key, "total_us", round(tot, 1), {k: round(v / tot, 3) for k, v in cls.items()}, "igemm bytes/launch", (rd + wr) / len(ig))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
