"""Profiling targets for ncu (run under gpurun; see profiles/README.md).
  python tools/profile_target.py unet   -> one 32-forward U-Net micro-batch @64x64 between cudaProfilerStart/Stop
                                           (eager replay, DM_GRAPH=0), preceded by an identical warm-up
  python tools/profile_target.py ops    -> the three dominant kernel shapes launched alone through the op-level ABI
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DM_GRAPH", "0")


def unet():
    from diff_mining_b200.engine import Engine
    from oracle import sd15

    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    eng = Engine(0)
    eng.load_state_dict(usd, "unet.")
    eng.finalize()
    g = torch.Generator().manual_seed(5)
    for i in range(2):
        eng.set_context(i, torch.randn(77, 768, generator=g))
    Bf = int(os.environ.get("DM_BF", "32"))
    x = torch.randn(Bf, 4, 64, 64, generator=g).cuda()
    t = torch.randint(100, 700, (Bf,), generator=g).cuda()
    slots = [i % 2 for i in range(Bf)]
    for _ in range(2):
        eng.unet_eps(x, t, slots)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.unet_eps(x, t, slots)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def ops():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gpu_optest as o

    lib, ptr, stream = o.lib, o.ptr, o.stream

    def conv(N, H, W, C0, Cout, ks, geglu=0, res=False):
        x = torch.randn(N, H, W, C0, device="cuda").half()
        w = (torch.randn(Cout, ks * ks * C0, device="cuda") / (ks * ks * C0) ** 0.5).half()
        b = torch.randn(Cout, device="cuda")
        r = torch.randn(N * H * W, Cout, device="cuda").half() if res else None
        out = torch.empty(N * H * W, Cout // 2 if geglu else Cout, device="cuda", dtype=torch.float16)
        return lambda: o.check(lib.dm_op_conv(ptr(x), None, N, H, W, C0, 0, ptr(w), Cout, ks, 1, 0, ptr(b), None, ptr(r), ptr(out), 0, geglu, 0, 0, stream())), (x, w, b, r, out)

    def attn(B, T, D):
        C = 8 * D
        qkv = torch.randn(B, T, 3 * C, device="cuda").half()
        out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
        args = (ptr(qkv[..., :C]), ptr(qkv[..., C:2 * C]), ptr(qkv[..., 2 * C:]), 3 * C, 3 * C, 3 * C, T * 3 * C, T * 3 * C, T * 3 * C, B, 8, D, T, T, 0, None, ptr(out), C, stream())
        return lambda: o.check(lib.dm_op_attention(*args)), (qkv, out)

    def xattn(B, T, D):
        C = 8 * D
        q = torch.randn(B, T, C, device="cuda").half()
        kv = torch.randn(2, 77, 2 * C, device="cuda").half()
        idx = torch.arange(B, device="cuda", dtype=torch.int32) % 2
        out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
        args = (ptr(q), ptr(kv[..., :C]), ptr(kv[..., C:]), C, 2 * C, 2 * C, T * C, 77 * 2 * C, 77 * 2 * C, B, 8, D, T, 77, 2, ptr(idx), ptr(out), C, stream())
        return lambda: o.check(lib.dm_op_attention(*args)), (q, kv, idx, out)

    def gn(N, HW, C):
        x = torch.randn(N, HW, C, device="cuda").half()
        g = torch.randn(C, device="cuda")
        b = torch.randn(C, device="cuda")
        out = torch.empty_like(x)
        return lambda: o.check(lib.dm_op_groupnorm(ptr(x), None, N, HW, C, 0, ptr(g), ptr(b), 1e-5, 1, ptr(out), stream())), (x, g, b, out)

    def conv_gn(N, H, C, silu):
        """3x3 conv with GroupNorm statistics in its epilogue + the fold/apply GroupNorm kernel (two launches)"""
        import ctypes
        P, I = ctypes.c_void_p, ctypes.c_int
        lib.dm_op_conv_gn.argtypes = [P, I, I, I, I, P, I, I, P, P, P, P, P, ctypes.c_float, I, P, P, P]
        x = torch.randn(N, H, H, C, device="cuda").half()
        w = (torch.randn(C, 9 * C, device="cuda") / (9 * C) ** 0.5).half()
        b, g, bt = torch.randn(C, device="cuda"), torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
        out = torch.empty(N * H * H, C, device="cuda", dtype=torch.float16)
        gno = torch.empty_like(out)
        return lambda: o.check(lib.dm_op_conv_gn(ptr(x), N, H, H, C, ptr(w), C, 3, ptr(b), None, None, ptr(g), ptr(bt), 1e-5, silu,
                                                 ptr(out), ptr(gno), stream())), (x, w, b, g, bt, out, gno)

    def ln(rows, C):
        x = torch.randn(rows, C, device="cuda").half()
        g = torch.randn(C, device="cuda")
        b = torch.randn(C, device="cuda")
        out = torch.empty_like(x)
        return lambda: o.check(lib.dm_op_layernorm(ptr(x), rows, C, ptr(g), ptr(b), 1e-5, ptr(out), stream())), (x, g, b, out)

    targets = [conv(32, 64, 64, 320, 320, 3), conv(32, 64, 64, 320, 2560, 1, geglu=1), conv(32, 64, 64, 320, 320, 1, res=True),
               conv(32, 16, 16, 1280, 1280, 3), conv(32, 32, 32, 1280, 1280, 3), attn(32, 4096, 40), attn(32, 1024, 80),
               xattn(32, 4096, 40), gn(32, 4096, 320), ln(32 * 4096, 320), conv_gn(32, 64, 320, 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for f, _ in targets:
        for _ in range(2):
            f()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for f, _ in targets:
        flush.zero_()
        f()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def attn():
    """self-attention d=40 T=4096 alone (B from DM_BF, default 8), for ncu --set full"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gpu_optest as o

    lib, ptr, stream = o.lib, o.ptr, o.stream
    B, T, D = int(os.environ.get("DM_BF", "8")), 4096, 40
    C = 8 * D
    qkv = torch.randn(B, T, 3 * C, device="cuda").half()
    out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
    args = (ptr(qkv[..., :C]), ptr(qkv[..., C:2 * C]), ptr(qkv[..., 2 * C:]), 3 * C, 3 * C, 3 * C, T * 3 * C, T * 3 * C, T * 3 * C, B, 8, D, T, T, 0, None, ptr(out), C, stream())
    for _ in range(2):
        o.check(lib.dm_op_attention(*args))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    o.check(lib.dm_op_attention(*args))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def layers():
    """per-layer CUDA-event table of one micro-batch (DMPROF lines on stdout)"""
    from diff_mining_b200.engine import Engine
    from oracle import sd15

    os.environ["DM_PROFILE_VERBOSE"] = "1"
    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    eng = Engine(0)
    eng.load_state_dict(usd, "unet.")
    if os.environ.get("DM_PROFILE_KIND") == "vae":  # DM_LAT = image size then
        eng.load_state_dict(sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1), "vae.")
    eng.finalize()
    g = torch.Generator().manual_seed(5)
    for i in range(2):
        eng.set_context(i, torch.randn(77, 768, generator=g))
    Bf = int(os.environ.get("DM_BF", "32"))
    lat = int(os.environ.get("DM_LAT", "64"))
    pr = eng.profile_unet(Bf, lat, lat, iters=5)
    print("DMPROF_TOTAL", pr)


def _engine(vae=False):
    from diff_mining_b200.engine import Engine
    from oracle import sd15

    eng = Engine(0)
    eng.load_state_dict(sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0), "unet.")
    if vae:
        eng.load_state_dict(sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1), "vae.")
    eng.finalize()
    g = torch.Generator().manual_seed(5)
    for i in range(15):
        eng.set_context(i, torch.randn(77, 768, generator=g))
    return eng, g


def _bracket(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def typ():
    """ONE micro-batch of the benched config-2 plan: 54 forwards @64x64 as 27 (eps,t) draws x {c, uncond}, shared prefix"""
    eng, g = _engine()
    x0 = torch.randn(1, 4, 64, 64, generator=g)
    noise = torch.randn(27, 4, 64, 64, generator=g)
    t = torch.randint(100, 700, (27,), generator=g)
    _bracket(lambda: eng.typicality(x0, noise, t, [1, 0], max_forwards=56))


def typ5():
    """ONE micro-batch of the benched config-5 plan: 15 forwards @128x128 = one (eps,t) draw x (14 conditions + uncond)"""
    eng, g = _engine()
    x0 = torch.randn(1, 4, 128, 128, generator=g)
    noise = torch.randn(1, 4, 128, 128, generator=g)
    t = torch.randint(0, 1000, (1,), generator=g)
    _bracket(lambda: eng.typicality(x0, noise, t, list(range(1, 15)) + [0]))


def dift():
    """ONE micro-batch of the benched config-3 plan: 64 DIFT members @64x64, t=161, up_ft_index=1"""
    eng, g = _engine()
    lat = torch.randn(64, 4, 64, 64, generator=g)
    nz = torch.randn(64, 4, 64, 64, generator=g)
    _bracket(lambda: eng.dift(lat, nz, 161, 1, 8, up_ft_index=1))


def vae():
    """VAE encode of 16 images 512x512 (config 2's per-step encode)"""
    eng, g = _engine(vae=True)
    img = torch.rand(16, 3, 512, 512, generator=g) * 2 - 1
    _bracket(lambda: eng.vae_encode(img, None))


if __name__ == "__main__":
    {"unet": unet, "ops": ops, "layers": layers, "attn": attn, "typ": typ, "typ5": typ5, "dift": dift, "vae": vae}[sys.argv[1]]()
