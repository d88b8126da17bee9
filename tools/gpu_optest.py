"""Quick operator-level checks of libdm_b200.so against torch ops on the GPU box (debug helper; the
pytest versions live in tests/test_gpu_ops.py).  Run:  python tools/gpu_optest.py [conv|attn|norm|all]"""
import ctypes
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.environ.get("DM_LIB") or os.path.join(ROOT, "diff-mining_b200", "libdm_b200.so"))
lib.dm_last_error.restype = ctypes.c_char_p
P = ctypes.c_void_p
I = ctypes.c_int
L = ctypes.c_int64
lib.dm_op_conv.argtypes = [P, P, I, I, I, I, I, P, I, I, I, I, P, P, P, P, I, I, I, I, P]
lib.dm_op_attention.argtypes = [P, P, P, L, L, L, L, L, L, I, I, I, I, I, I, P, P, L, P]
lib.dm_op_groupnorm.argtypes = [P, P, I, I, I, I, P, P, ctypes.c_float, I, P, P]
lib.dm_op_layernorm.argtypes = [P, L, I, P, P, ctypes.c_float, P, P]


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def check(rc):
    if rc != 0:
        raise RuntimeError(lib.dm_last_error().decode())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_w(w):  # OIHW -> [O, kh*kw*I] tap-major
    O, Ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(O, kh * kw * Ci).contiguous()


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    ok = err <= tol * scale and math.isfinite(err)
    print(f"{'PASS' if ok else 'FAIL'} {name}: max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e}", flush=True)
    return ok


def conv_case(N, H, W, C0, C1, Cout, ks, stride=1, vae_pad=0, bias=True, rowbias=False, residual=False, geglu=False,
              silu=False, out_f32=False, bn=0):
    g = torch.Generator(device="cuda").manual_seed(1234 + N + H + C0 + Cout)
    x = torch.randn(N, H, W, C0, device="cuda", generator=g).half()
    x2 = torch.randn(N, H, W, C1, device="cuda", generator=g).half() if C1 else None
    Cin = C0 + C1
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / math.sqrt(Cin * ks * ks)).half()
    b = torch.randn(Cout, device="cuda", generator=g).half().float() if bias else None
    xin = x if x2 is None else torch.cat([x, x2], dim=-1)
    xn = xin.permute(0, 3, 1, 2).float()
    if stride == 2 and vae_pad:
        xn = F.pad(xn, (0, 1, 0, 1))
        ref = F.conv2d(xn, w.float(), b, stride=2, padding=0)
    else:
        ref = F.conv2d(xn, w.float(), b, stride=stride, padding=ks // 2)
    Ho, Wo = ref.shape[2], ref.shape[3]
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    rb = rs = None
    if rowbias:
        rb = torch.randn(N, Cout, device="cuda", generator=g).half()
        ref = ref.half().float() + rb.float().repeat_interleave(Ho * Wo, dim=0)
    if silu:
        ref = F.silu(ref.half().float())
    if geglu:
        r16 = ref.half().float()
        ref = r16[:, 0::2] * F.gelu(r16[:, 1::2])
    if residual:
        rs = torch.randn(N * Ho * Wo, Cout, device="cuda", generator=g).half()
        ref = ref.half().float() + rs.float()
    ocols = Cout // 2 if geglu else Cout
    out = torch.full((N * Ho * Wo, ocols), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.float16)
    wp = pack_w(w)
    check(lib.dm_op_conv(ptr(x), ptr(x2), N, H, W, C0, C1, ptr(wp), Cout, ks, stride, vae_pad, ptr(b), ptr(rb), ptr(rs),
                         ptr(out), int(out_f32), int(geglu), int(silu), bn, stream()))
    torch.cuda.synchronize()
    name = f"conv N{N} {H}x{W} C{C0}+{C1}->{Cout} k{ks} s{stride} vp{vae_pad} rb{int(rowbias)} res{int(residual)} geglu{int(geglu)} silu{int(silu)} f32{int(out_f32)} bn{bn}"
    return report(name, out, ref, tol=1e-2)


def attn_case(B, T, Tk, D, cross=False, heads=8):
    g = torch.Generator(device="cuda").manual_seed(99 + B + T + D)
    C = heads * D
    if not cross:
        qkv = torch.randn(B, T, 3 * C, device="cuda", generator=g).half()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        ldq = ldk = ldv = 3 * C
        bsq = bsk = bsv = T * 3 * C
        kvb = 0
        kvi = None
        kk, vv = k, v
    else:
        nslots = 3
        q = torch.randn(B, T, C, device="cuda", generator=g).half()
        kc = torch.randn(nslots, Tk, C, device="cuda", generator=g).half()
        vc = torch.randn(nslots, Tk, C, device="cuda", generator=g).half()
        kvi = (torch.arange(B, device="cuda", dtype=torch.int32) % nslots).contiguous()
        kk, vv = kc[kvi.long()], vc[kvi.long()]
        k, v = kc, vc
        ldq, ldk, ldv = C, C, C
        bsq, bsk, bsv = T * C, Tk * C, Tk * C
        kvb = nslots
    out = torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.float16)
    check(lib.dm_op_attention(ptr(q), ptr(k), ptr(v), ldq, ldk, ldv, bsq, bsk, bsv, B, heads, D, T, Tk, kvb, ptr(kvi),
                              ptr(out), C, stream()))
    torch.cuda.synchronize()
    qh = q.float().reshape(B, T, heads, D).transpose(1, 2)
    kh = kk.float().reshape(B, Tk, heads, D).transpose(1, 2)
    vh = vv.float().reshape(B, Tk, heads, D).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, T, C)
    return report(f"attn B{B} Tq{T} Tk{Tk} D{D} cross{int(cross)}", out, ref, tol=5e-3)


def norm_cases():
    ok = True
    g = torch.Generator(device="cuda").manual_seed(7)
    for (N, HW, C0, C1, silu, eps) in [(2, 4096, 320, 0, 1, 1e-5), (3, 1024, 640, 320, 1, 1e-5), (2, 256, 1280, 1280, 0, 1e-6),
                                        (1, 64, 1280, 640, 1, 1e-5), (2, 100, 128, 0, 1, 1e-6), (5, 4096, 640, 320, 1, 1e-5)]:
        C = C0 + C1
        x = (torch.randn(N, HW, C0, device="cuda", generator=g) * 1.5 + 0.3).half()
        x2 = (torch.randn(N, HW, C1, device="cuda", generator=g) * 0.7 - 0.2).half() if C1 else None
        gamma = torch.randn(C, device="cuda", generator=g)
        beta = torch.randn(C, device="cuda", generator=g)
        out = torch.full((N, HW, C), float("nan"), device="cuda", dtype=torch.float16)
        check(lib.dm_op_groupnorm(ptr(x), ptr(x2), N, HW, C0, C1, ptr(gamma), ptr(beta), eps, silu, ptr(out), stream()))
        xin = x if x2 is None else torch.cat([x, x2], -1)
        ref = F.group_norm(xin.float().permute(0, 2, 1), 32, gamma, beta, eps)
        if silu:
            ref = F.silu(ref)
        ref = ref.permute(0, 2, 1)
        ok &= report(f"groupnorm N{N} HW{HW} C{C0}+{C1} silu{silu}", out, ref, tol=2e-3)
    for (rows, C) in [(4096, 320), (1000, 640), (77, 1280)]:
        x = (torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5).half()
        gamma = torch.randn(C, device="cuda", generator=g)
        beta = torch.randn(C, device="cuda", generator=g)
        out = torch.full((rows, C), float("nan"), device="cuda", dtype=torch.float16)
        check(lib.dm_op_layernorm(ptr(x), rows, C, ptr(gamma), ptr(beta), 1e-5, ptr(out), stream()))
        ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
        ok &= report(f"layernorm rows{rows} C{C}", out, ref, tol=2e-3)
    return ok


def conv_cases():
    ok = True
    ok &= conv_case(1, 1, 128, 64, 0, 64, 1, bias=False)            # smallest plain GEMM: one tile, one k-iter
    ok &= conv_case(1, 1, 256, 128, 0, 128, 1)                      # 2 m-tiles, 2 k-iters
    ok &= conv_case(1, 1, 77, 768, 0, 320, 1)                       # ragged M, BN=160
    ok &= conv_case(2, 16, 16, 64, 0, 64, 3, bias=False)            # 3x3 halo via TMA OOB
    ok &= conv_case(2, 64, 64, 320, 0, 320, 3, rowbias=True)        # the dominant ResNet conv
    ok &= conv_case(2, 32, 32, 640, 320, 640, 3, residual=True)     # concat sources
    ok &= conv_case(4, 8, 8, 1280, 1280, 1280, 3)                   # multi-image tiles
    ok &= conv_case(2, 17, 23, 128, 0, 256, 3)                      # ragged spatial
    ok &= conv_case(2, 32, 32, 320, 0, 320, 3, stride=2)            # U-Net downsample
    ok &= conv_case(2, 33, 31, 128, 0, 128, 3, stride=2)            # odd sizes
    ok &= conv_case(2, 32, 32, 128, 0, 128, 3, stride=2, vae_pad=1)  # VAE downsample
    ok &= conv_case(1, 33, 31, 128, 0, 128, 3, stride=2, vae_pad=1)
    ok &= conv_case(1, 64, 64, 320, 0, 2560, 1, geglu=True)         # GEGLU
    ok &= conv_case(1, 1, 32, 320, 0, 1280, 1, silu=True)           # time MLP
    ok &= conv_case(2, 16, 16, 320, 0, 16, 3)                       # conv_out (N=16)
    ok &= conv_case(1, 32, 32, 512, 0, 1024, 1, out_f32=True, bias=False)
    ok &= conv_case(2, 16, 16, 1280, 640, 1280, 1)                  # shortcut 1x1 over concat
    for bn in (16, 32, 64, 128, 160, 256):
        ok &= conv_case(1, 16, 32, 128, 0, 320 if bn == 160 else 256, 3, bn=bn)
    return ok


def attn_cases():
    ok = True
    ok &= attn_case(1, 128, 128, 40)
    ok &= attn_case(2, 4096, 4096, 40)
    ok &= attn_case(2, 1024, 1024, 80)
    ok &= attn_case(2, 256, 256, 160)
    ok &= attn_case(3, 64, 64, 160)
    ok &= attn_case(2, 1000, 1000, 40)
    ok &= attn_case(1, 300, 300, 80)
    ok &= attn_case(3, 257, 257, 40)
    ok &= attn_case(1, 129, 129, 40)
    ok &= attn_case(4, 4096, 77, 40, cross=True)
    ok &= attn_case(4, 1024, 77, 80, cross=True)
    ok &= attn_case(4, 256, 77, 160, cross=True)
    ok &= attn_case(5, 64, 77, 160, cross=True)
    return ok


def bench_conv():
    # rough timing of the dominant shapes
    for (N, H, W, C0, Cout, ks) in [(16, 64, 64, 320, 320, 3), (16, 32, 32, 640, 640, 3), (16, 16, 16, 1280, 1280, 3),
                                     (16, 64, 64, 320, 2560, 1), (16, 64, 64, 1280, 320, 1)]:
        x = torch.randn(N, H, W, C0, device="cuda").half()
        w = torch.randn(Cout, ks * ks * C0, device="cuda").half()
        out = torch.empty(N * H * W, Cout, device="cuda", dtype=torch.float16)
        for _ in range(3):
            check(lib.dm_op_conv(ptr(x), None, N, H, W, C0, 0, ptr(w), Cout, ks, 1, 0, None, None, None, ptr(out), 0, 0, 0, 0, stream()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 10
        for _ in range(iters):
            check(lib.dm_op_conv(ptr(x), None, N, H, W, C0, 0, ptr(w), Cout, ks, 1, 0, None, None, None, ptr(out), 0, 0, 0, 0, stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * N * H * W * Cout * ks * ks * C0
        print(f"BENCH conv N{N} {H}x{W} {C0}->{Cout} k{ks}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    for (B, T, D) in [(16, 4096, 40), (16, 1024, 80), (16, 256, 160)]:
        C = 8 * D
        qkv = torch.randn(B, T, 3 * C, device="cuda").half()
        out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
        args = (ptr(qkv[..., :C]), ptr(qkv[..., C:2 * C]), ptr(qkv[..., 2 * C:]), 3 * C, 3 * C, 3 * C, T * 3 * C, T * 3 * C,
                T * 3 * C, B, 8, D, T, T, 0, None, ptr(out), C, stream())
        for _ in range(3):
            check(lib.dm_op_attention(*args))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            check(lib.dm_op_attention(*args))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * B * 8 * T * T * D
        print(f"BENCH attn B{B} T{T} D{D}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def bench_attn():
    """event-timed attention shapes of one 32-forward micro-batch (self + cross)"""
    for (B, T, Tk, D) in [(32, 4096, 4096, 40), (32, 1024, 1024, 80), (32, 256, 256, 160), (32, 4096, 77, 40), (32, 1024, 77, 80)]:
        C = 8 * D
        if Tk == T:
            qkv = torch.randn(B, T, 3 * C, device="cuda").half()
            out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
            args = (ptr(qkv[..., :C]), ptr(qkv[..., C:2 * C]), ptr(qkv[..., 2 * C:]), 3 * C, 3 * C, 3 * C, T * 3 * C, T * 3 * C,
                    T * 3 * C, B, 8, D, T, T, 0, None, ptr(out), C, stream())
        else:
            q = torch.randn(B, T, C, device="cuda").half()
            kv = torch.randn(2, Tk, 2 * C, device="cuda").half()
            kvi = (torch.arange(B, device="cuda", dtype=torch.int32) % 2).contiguous()
            out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
            args = (ptr(q), ptr(kv[..., :C]), ptr(kv[..., C:]), C, 2 * C, 2 * C, T * C, Tk * 2 * C, Tk * 2 * C, B, 8, D, T, Tk, 2,
                    ptr(kvi), ptr(out), C, stream())
        for _ in range(3):
            check(lib.dm_op_attention(*args))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            check(lib.dm_op_attention(*args))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * B * 8 * T * Tk * D
        print(f"BENCH attn B{B} T{T} Tk{Tk} D{D} DM_ATTN2={os.environ.get('DM_ATTN2','1')}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    t0 = time.time()
    ok = True
    if which in ("norm", "all"):
        ok &= norm_cases()
    if which in ("conv", "all"):
        ok &= conv_cases()
    if which in ("attn", "all"):
        ok &= attn_cases()
    if which in ("bench", "all"):
        bench_conv()
    if which in ("benchattn",):
        bench_attn()
    print("ALL OK" if ok else "SOME FAILED", f"({time.time()-t0:.1f}s)")
    sys.exit(0 if ok else 1)
