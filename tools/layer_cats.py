"""igemm time by layer family from `tools/profile_target.py layers` logs: python tools/layer_cats.py <log> [<log> ...]"""
import collections
import re
import sys


def cat(n):
    if "ff.net.0" in n:
        return "geglu"
    if "ff.net.2" in n:
        return "ff2"
    for k in ["proj_out", "proj_in", "to_out", "to_q", "qkv", "conv_shortcut", "conv_in", "conv_out", "downsamplers", "upsamplers"]:
        if k in n:
            return k
    if n.endswith("conv1") or n.endswith("conv2"):
        return "conv3x3"
    return "misc"


for f in sys.argv[1:]:
    c = collections.defaultdict(float)
    for ln in open(f):
        m = re.match(r"DMPROF (\S+)\s+cls=(\d) ms=([\d.]+) gflop=([\d.]+)", ln)
        if m and m.group(2) == "0":
            c[cat(m.group(1))] += float(m.group(3))
    print(f.split("/")[-1], " ".join(f"{a} {b:.3f}" for a, b in sorted(c.items())), f"| igemm {sum(c.values()):.3f}")
