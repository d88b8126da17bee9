"""Debug: clock64 timeline of one CTA of the warp-specialised attention kernel (DM_ATTN_TRACE=1).
   python tools/attn_trace.py [D] [T]  -> ATTNTRACE lines on stderr (see abi_ops.cu)"""
import os
import sys

os.environ["DM_ATTN_TRACE"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_optest as o

D = int(sys.argv[1]) if len(sys.argv) > 1 else 40
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
B, C = 8, 8 * D
qkv = torch.randn(B, T, 3 * C, device="cuda").half()
out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
args = (o.ptr(qkv[..., :C]), o.ptr(qkv[..., C:2 * C]), o.ptr(qkv[..., 2 * C:]), 3 * C, 3 * C, 3 * C, T * 3 * C, T * 3 * C, T * 3 * C,
        B, 8, D, T, T, 0, None, o.ptr(out), C, o.stream())
o.check(o.lib.dm_op_attention(*args))
torch.cuda.synchronize()
