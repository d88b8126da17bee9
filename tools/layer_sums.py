"""Sum a `tools/profile_target.py layers` log by kernel family: usage  python tools/layer_sums.py <log> [label]"""
import re
import sys

rows = []
for ln in open(sys.argv[1]):
    m = re.match(r"DMPROF (\S+)\s+cls=(\d) ms=([\d.]+) gflop=([\d.]+)", ln)
    if m:
        rows.append((m.group(1), int(m.group(2)), float(m.group(3))))
tot = {0: 0.0, 1: 0.0, 2: 0.0}
fam = {"ln": 0.0, "gn_apply": 0.0, "gn_full": 0.0, "conv3x3": 0.0}
for n, c, ms in rows:
    tot[c] += ms
    if c == 2 and "transformer_blocks" in n and "norm" in n:
        fam["ln"] += ms
    elif c == 2 and n.endswith(".apply"):
        fam["gn_apply"] += ms
    elif c == 2 and "norm" in n:
        fam["gn_full"] += ms
    elif c == 0 and (n.endswith("conv1") or n.endswith("conv2")):
        fam["conv3x3"] += ms
label = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
print(f"{label}: total {sum(tot.values()):.3f} igemm {tot[0]:.3f} attn {tot[1]:.3f} other {tot[2]:.3f} | "
      + " ".join(f"{k} {v:.3f}" for k, v in fam.items()))
