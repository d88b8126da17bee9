"""GPU parity tests, operator level: every hand-written kernel called through the C ABI (dm_op_*) against the same
op in plain PyTorch fp32 on identical seeded inputs.  fp16 outputs => tolerance = a few fp16 ulps of the output scale
(the kernels accumulate in fp32 like the reference's cuDNN/cuBLAS/xformers kernels do)."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from diff_mining_b200 import _abi

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _abi.load()


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(lib, rc):
    assert rc == 0, lib.dm_last_error().decode()


def pack_w(w):  # OIHW -> [O, kh*kw*I] tap-major (the layout dm_op_conv documents)
    O, Ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(O, kh * kw * Ci).contiguous()


def max_rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-9)).item()


CONV_CASES = [
    # N, H, W, C0, C1, Cout, ks, stride, vae_pad, rowbias, residual, geglu, silu, f32, bn
    (1, 1, 128, 64, 0, 64, 1, 1, 0, 0, 0, 0, 0, 0, 0),       # one tile, one k-chunk
    (1, 1, 77, 768, 0, 320, 1, 1, 0, 0, 0, 0, 0, 0, 0),      # ragged M (context K/V projection)
    (2, 16, 16, 64, 0, 64, 3, 1, 0, 0, 0, 0, 0, 0, 0),       # halo through TMA out-of-bounds fill
    (2, 64, 64, 320, 0, 320, 3, 1, 0, 1, 0, 0, 0, 0, 0),     # dominant ResNet conv + time-embedding rowbias
    (2, 32, 32, 640, 320, 640, 3, 1, 0, 0, 1, 0, 0, 0, 0),   # cat(h, skip) sources + residual
    (4, 8, 8, 1280, 1280, 1280, 3, 1, 0, 0, 0, 0, 0, 0, 0),  # several images per M tile
    (2, 17, 23, 128, 0, 256, 3, 1, 0, 0, 0, 0, 0, 0, 0),     # ragged spatial extent
    (1, 1, 1, 64, 0, 64, 3, 1, 0, 0, 0, 0, 0, 0, 0),         # 1x1 image: everything but the centre tap is padding
    (2, 32, 32, 320, 0, 320, 3, 2, 0, 0, 0, 0, 0, 0, 0),     # U-Net downsample (stride 2, pad 1)
    (2, 33, 31, 128, 0, 128, 3, 2, 0, 0, 0, 0, 0, 0, 0),     # odd sizes -> ceil
    (2, 32, 32, 128, 0, 128, 3, 2, 1, 0, 0, 0, 0, 0, 0),     # VAE downsample (pad (0,1,0,1))
    (1, 33, 31, 128, 0, 128, 3, 2, 1, 0, 0, 0, 0, 0, 0),     # odd sizes -> floor
    (1, 64, 64, 320, 0, 2560, 1, 1, 0, 0, 0, 1, 0, 0, 0),    # GEGLU epilogue
    (1, 1, 32, 320, 0, 1280, 1, 1, 0, 0, 0, 0, 1, 0, 0),     # time MLP (SiLU epilogue)
    (2, 16, 16, 320, 0, 16, 3, 1, 0, 0, 0, 0, 0, 0, 0),      # conv_out (N = 16)
    (1, 32, 32, 512, 0, 1024, 1, 1, 0, 0, 0, 0, 0, 1, 0),    # fp32 output (VAE attention scores)
    (2, 16, 16, 1280, 640, 1280, 1, 1, 0, 0, 0, 0, 0, 0, 0),  # shortcut 1x1 over cat
] + [(1, 16, 32, 128, 0, 320 if bn == 160 else 256, 3, 1, 0, 0, 0, 0, 0, 0, bn) for bn in (16, 32, 64, 128, 160, 256)]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_igemm(lib, case):
    N, H, W, C0, C1, Cout, ks, stride, vae_pad, rowbias, residual, geglu, silu, f32, bn = case
    g = torch.Generator(device="cuda").manual_seed(1234 + N + H + C0 + Cout)
    x = torch.randn(N, H, W, C0, device="cuda", generator=g).half()
    x2 = torch.randn(N, H, W, C1, device="cuda", generator=g).half() if C1 else None
    Cin = C0 + C1
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / math.sqrt(Cin * ks * ks)).half()
    b = torch.randn(Cout, device="cuda", generator=g).half().float()
    xin = x if x2 is None else torch.cat([x, x2], dim=-1)
    xn = xin.permute(0, 3, 1, 2).float()
    if stride == 2 and vae_pad:
        ref = F.conv2d(F.pad(xn, (0, 1, 0, 1)), w.float(), b, stride=2, padding=0)
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
    out = torch.full((N * Ho * Wo, ocols), float("nan"), device="cuda", dtype=torch.float32 if f32 else torch.float16)
    check(lib, lib.dm_op_conv(ptr(x), ptr(x2), N, H, W, C0, C1, ptr(pack_w(w)), Cout, ks, stride, vae_pad, ptr(b), ptr(rb),
                              ptr(rs), ptr(out), f32, geglu, silu, bn, stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert max_rel(out, ref) < (1e-5 if f32 else 2.5e-3)  # fp16 output: ~2 ulp of the output scale


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[5] >= 128 and not c[13] and c[14] in (0, 128, 160, 256)],
                         ids=lambda c: "-".join(map(str, c)))
def test_conv_igemm_cta_pairs(lib, case):
    """same cases on the cta_group::2 path (256-row tiles over a CTA pair), forced even for tiny problems: odd M-tile
    counts (dummy second tile), ragged tiles, every epilogue option"""
    check(lib, lib.dm_op_set_variant(b"igemm_pair", 2))
    try:
        test_conv_igemm(lib, case)
    finally:
        check(lib, lib.dm_op_set_variant(b"igemm_pair", -1))


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[5] >= 160 and not c[13] and c[14] in (0, 160, 256)],
                         ids=lambda c: "-".join(map(str, c)))
def test_conv_igemm_four_epilogue_groups(lib, case):
    """same cases with four epilogue warpgroups (the short-K variant), forced for every eligible tile shape"""
    check(lib, lib.dm_op_set_variant(b"igemm_pair", 0))
    check(lib, lib.dm_op_set_variant(b"igemm_ng4", 2))
    try:
        test_conv_igemm(lib, case)
    finally:
        check(lib, lib.dm_op_set_variant(b"igemm_ng4", -1))
        check(lib, lib.dm_op_set_variant(b"igemm_pair", -1))


WS_CASES = [
    # N, H, W, C0, C1, Cout, ks, stride, vae_pad, rowbias, residual, geglu, silu, f32, bn
    (3, 64, 64, 320, 0, 320, 1, 1, 0, 0, 1, 0, 0, 0, 0),     # attn.to_out / proj_out: two N-tiles of 160, residual
    (2, 64, 64, 320, 0, 2560, 1, 1, 0, 0, 0, 1, 0, 0, 0),    # GEGLU: ten N-tiles of 256 (four epilogue groups)
    (3, 64, 64, 320, 0, 960, 1, 1, 0, 0, 0, 0, 0, 0, 0),     # qkv: six N-tiles, no bias epilogue options
    (1, 37, 41, 320, 0, 320, 1, 1, 0, 0, 1, 0, 0, 0, 0),     # ragged M, odd number of M-tiles (dummy second tile of the last pair)
    (1, 16, 16, 64, 0, 640, 1, 1, 0, 0, 0, 0, 1, 0, 0),      # one k-chunk, fewer M-units than CTA pairs, SiLU
    (2, 16, 16, 256, 0, 256, 1, 1, 0, 1, 1, 0, 0, 0, 0),     # one N-tile of 256, rowbias + residual
    (5, 32, 32, 320, 0, 320, 3, 1, 0, 1, 0, 0, 0, 0, 0),     # 3x3 (K = 2880): must fall back to the streaming-B kernel
]


@pytest.mark.parametrize("case", WS_CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_igemm_weight_stationary(lib, case):
    """short-K Linears on weight-stationary CTA pairs (igemm.cuh: WS -- the B tile of one N-tile stays in shared memory
    while the pair walks down M), forced for every eligible shape; bit-identical to the streaming-B kernel"""
    N, H, W, C0, C1, Cout, ks, stride, vae_pad, rowbias, residual, geglu, silu, f32, bn = case
    check(lib, lib.dm_op_set_variant(b"igemm_ws", 2))
    try:
        test_conv_igemm(lib, case)
    finally:
        check(lib, lib.dm_op_set_variant(b"igemm_ws", -1))
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, H, W, C0, device="cuda", generator=g).half()
    w = (torch.randn(Cout, C0, ks, ks, device="cuda", generator=g) / math.sqrt(C0 * ks * ks)).half()
    b = torch.randn(Cout, device="cuda", generator=g)
    rs = torch.randn(N * H * W, Cout, device="cuda", generator=g).half() if residual else None
    outs = []
    for mode in (2, 0):
        check(lib, lib.dm_op_set_variant(b"igemm_ws", mode))
        out = torch.full((N * H * W, Cout // 2 if geglu else Cout), float("nan"), device="cuda", dtype=torch.float16)
        check(lib, lib.dm_op_conv(ptr(x), None, N, H, W, C0, 0, ptr(pack_w(w)), Cout, ks, 1, 0, ptr(b), None, ptr(rs), ptr(out), 0,
                                  geglu, silu, 0, stream()))
        torch.cuda.synchronize()
        outs.append(out)
    check(lib, lib.dm_op_set_variant(b"igemm_ws", -1))
    assert torch.equal(outs[0], outs[1])


def test_groupnorm_two_kernel_path_matches_fused(lib):
    """the cluster-fused GroupNorm and the stats + apply path agree (both deterministic)"""
    g = torch.Generator(device="cuda").manual_seed(7)
    N, HW, C = 3, 1024, 640
    x = torch.randn(N, HW, C, device="cuda", generator=g).half()
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    outs = []
    for mode in (1, 0):
        check(lib, lib.dm_op_set_variant(b"gn_fused", mode))
        out = torch.empty_like(x)
        check(lib, lib.dm_op_groupnorm(ptr(x), None, N, HW, C, 0, ptr(gamma), ptr(beta), 1e-5, 1, ptr(out), stream()))
        torch.cuda.synchronize()
        outs.append(out)
    check(lib, lib.dm_op_set_variant(b"gn_fused", -1))
    ref = F.silu(F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert max_rel(outs[0], ref) < 2e-3 and max_rel(outs[1], ref) < 2e-3
    assert max_rel(outs[0], outs[1]) < 1e-3


CONV_GN_CASES = [
    # N, H, W, Cin, Cout, ks, rowbias, residual, silu, bias mean, igemm_pair, igemm_ng4
    (3, 32, 32, 320, 320, 3, 1, 0, 1, 0.0, -1, -1),     # conv1 -> norm2 (+ time embedding)
    (2, 64, 64, 320, 320, 3, 0, 1, 0, 0.0, -1, -1),     # conv2 (+ residual) -> Transformer2DModel.norm, 32 tiles / image
    (5, 16, 16, 640, 1280, 3, 1, 0, 1, 0.0, -1, -1),    # two tiles per image, 1280 channels
    (2, 32, 32, 640, 640, 3, 1, 1, 1, 50.0, -1, -1),    # channel means >> std (cancellation-free sums)
    (3, 16, 32, 320, 640, 3, 1, 0, 1, 0.0, 2, -1),      # CTA pairs
    (2, 32, 32, 320, 320, 3, 0, 1, 1, 0.0, 0, 2),       # four epilogue warpgroups
    (2, 32, 32, 320, 320, 1, 0, 1, 0, 0.0, -1, -1),     # 1x1
    (3, 64, 64, 320, 320, 1, 0, 1, 1, 0.0, -1, -1),     # 1x1 with enough M-tiles for the weight-stationary pair kernel
    (2, 128, 64, 320, 320, 3, 1, 0, 1, 0.0, -1, -1),    # 8192 pixels: two fold + apply clusters per image
    (1, 128, 128, 320, 320, 3, 0, 1, 1, 0.0, -1, -1),   # 16384 pixels (config 5): four clusters per image
]


@pytest.mark.parametrize("case", CONV_GN_CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_with_groupnorm_statistics_in_the_epilogue(lib, case):
    """conv whose epilogue forms the GroupNorm statistics of its output + the fold/apply kernel (igemm.cuh: IgGn,
    norm.cuh: gn_fold_apply_kernel) against conv2d -> fp16 -> group_norm in fp32; the conv output itself must be
    bit-identical to the plain conv, and both outputs deterministic and independent of the batch an image is in"""
    N, H, W, Cin, Cout, ks, rowbias, residual, silu, bmean, pair, ng4 = case
    g = torch.Generator(device="cuda").manual_seed(99 + N + H + Cin + Cout)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / math.sqrt(Cin * ks * ks)).half()
    b = (torch.randn(Cout, device="cuda", generator=g) + bmean * torch.randn(Cout, device="cuda", generator=g).sign()).half().float()
    rb = torch.randn(N, Cout, device="cuda", generator=g).half() if rowbias else None
    rs = (torch.randn(N * H * W, Cout, device="cuda", generator=g) + 2.0).half() if residual else None
    gamma = torch.randn(Cout, device="cuda", generator=g)
    beta = torch.randn(Cout, device="cuda", generator=g)
    check(lib, lib.dm_op_set_variant(b"igemm_pair", pair))
    check(lib, lib.dm_op_set_variant(b"igemm_ng4", ng4))
    try:
        def run(n0, n1):
            n = n1 - n0
            out = torch.full((n * H * W, Cout), float("nan"), device="cuda", dtype=torch.float16)
            gn = torch.full_like(out, float("nan"))
            check(lib, lib.dm_op_conv_gn(ptr(x[n0:n1].contiguous()), n, H, W, Cin, ptr(pack_w(w)), Cout, ks, ptr(b),
                                         ptr(rb[n0:n1].contiguous()) if rowbias else None,
                                         ptr(rs[n0 * H * W:n1 * H * W].contiguous()) if residual else None,
                                         ptr(gamma), ptr(beta), 1e-5, silu, ptr(out), ptr(gn), stream()))
            torch.cuda.synchronize()
            return out, gn
        out, gn = run(0, N)
        out_b, gn_b = run(0, N)
        assert torch.equal(out, out_b) and torch.equal(gn, gn_b)
        plain = torch.full_like(out, float("nan"))
        check(lib, lib.dm_op_conv(ptr(x), None, N, H, W, Cin, 0, ptr(pack_w(w)), Cout, ks, 1, 0, ptr(b), ptr(rb), ptr(rs),
                                  ptr(plain), 0, 0, 0, 0, stream()))
        torch.cuda.synchronize()
        assert torch.equal(out, plain)
        ref = F.group_norm(out.float().view(N, H * W, Cout).permute(0, 2, 1), 32, gamma, beta, 1e-5)
        ref = (F.silu(ref) if silu else ref).permute(0, 2, 1).reshape(N * H * W, Cout)
        assert max_rel(gn, ref) < 2e-3
        out1, gn1 = run(N - 1, N)  # the last image alone
        assert torch.equal(gn1, gn[(N - 1) * H * W:])
    finally:
        check(lib, lib.dm_op_set_variant(b"igemm_ng4", -1))
        check(lib, lib.dm_op_set_variant(b"igemm_pair", -1))


ATTN_CASES = [
    # B, Tq, Tk, D, cross
    (1, 128, 128, 40, 0), (2, 4096, 4096, 40, 0), (2, 1024, 1024, 80, 0), (2, 256, 256, 160, 0), (3, 64, 64, 160, 0),
    (2, 1000, 1000, 40, 0), (1, 77, 77, 80, 0), (2, 130, 130, 160, 0),                      # ragged tiles
    (4, 4096, 77, 40, 1), (4, 1024, 77, 80, 1), (4, 256, 77, 160, 1), (5, 64, 77, 160, 1),  # text cross-attention
]


@pytest.mark.parametrize("case", ATTN_CASES, ids=lambda c: "-".join(map(str, c)))
def test_flash_attention(lib, case):
    B, T, Tk, D, cross = case
    heads = 8
    C = heads * D
    g = torch.Generator(device="cuda").manual_seed(99 + B + T + D)
    if not cross:
        qkv = torch.randn(B, T, 3 * C, device="cuda", generator=g).half()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        ld = (3 * C,) * 3
        bs = (T * 3 * C,) * 3
        kvb, kvi, kk, vv = 0, None, k, v
    else:
        nslots = 3
        q = torch.randn(B, T, C, device="cuda", generator=g).half()
        kv = torch.randn(nslots, Tk, 2 * C, device="cuda", generator=g).half()   # the engine's K|V cache layout
        k, v = kv[..., :C], kv[..., C:]
        kvi = (torch.arange(B, device="cuda", dtype=torch.int32) % nslots).contiguous()
        kk, vv = k[kvi.long()], v[kvi.long()]
        ld = (C, 2 * C, 2 * C)
        bs = (T * C, Tk * 2 * C, Tk * 2 * C)
        kvb = nslots
    out = torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.float16)
    check(lib, lib.dm_op_attention(ptr(q), ptr(k), ptr(v), ld[0], ld[1], ld[2], bs[0], bs[1], bs[2], B, heads, D, T, Tk, kvb,
                                   ptr(kvi), ptr(out), C, stream()))
    torch.cuda.synchronize()
    qh = q.float().reshape(B, T, heads, D).transpose(1, 2)
    kh = kk.float().reshape(B, Tk, heads, D).transpose(1, 2)
    vh = vv.float().reshape(B, Tk, heads, D).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, T, C)
    assert torch.isfinite(out).all()
    assert max_rel(out, ref) < 2e-3


@pytest.mark.parametrize("case", [c for c in ATTN_CASES if not c[4] and c[3] in (40, 80)] + [(20, 1024, 1024, 40, 0), (40, 300, 300, 80, 0)],
                         ids=lambda c: "-".join(map(str, c)))
def test_flash_attention_variants(lib, case):
    """`attn3` switch: 0 = one CTA per query block (attention2), 1 = persistent kernel with two softmax threads per query
    row, 2 = persistent kernel with one thread per row; the extra cases give a persistent CTA several work items"""
    for v in (0, 1, 2):
        check(lib, lib.dm_op_set_variant(b"attn3", v))
        try:
            test_flash_attention(lib, case)
        finally:
            check(lib, lib.dm_op_set_variant(b"attn3", -1))


XATTN_CASES = [
    # B, Tq, Tk, D: several items per persistent CTA, context-slot changes inside a CTA's range, ragged query blocks
    (54, 4096, 77, 40), (54, 1024, 77, 80), (7, 1000, 77, 40), (9, 300, 77, 80), (3, 128, 77, 40), (1, 77, 60, 80), (160, 256, 77, 40),
]


@pytest.mark.parametrize("case", XATTN_CASES, ids=lambda c: "-".join(map(str, c)))
def test_cross_attention_persistent_matches_one_shot(lib, case):
    """text cross-attention: the persistent kernel (one head and a contiguous range of (batch, query block) pairs per CTA, K / V kept while the
    context slot does not change, next Q prefetched) against SDPA and bit-identical to the one-CTA-per-block
    kernel (xattn = 1)"""
    B, T, Tk, D = case
    heads, nslots = 8, 5
    C = heads * D
    g = torch.Generator(device="cuda").manual_seed(7 + B + T + D)
    q = torch.randn(B, T, C, device="cuda", generator=g).half()
    kv = torch.randn(nslots, Tk, 2 * C, device="cuda", generator=g).half()
    kvi = torch.randint(0, nslots, (B,), device="cuda", generator=g, dtype=torch.int32)
    outs = []
    for mode in (3, 1):  # 3 = persistent kernel at head_dim 40 and 80 (the default, 2, uses it at head_dim 40 only)
        check(lib, lib.dm_op_set_variant(b"xattn", mode))
        out = torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.float16)
        check(lib, lib.dm_op_attention(ptr(q), ptr(kv[..., :C]), ptr(kv[..., C:]), C, 2 * C, 2 * C, T * C, Tk * 2 * C, Tk * 2 * C, B,
                                       heads, D, T, Tk, nslots, ptr(kvi), ptr(out), C, stream()))
        torch.cuda.synchronize()
        outs.append(out)
    check(lib, lib.dm_op_set_variant(b"xattn", -1))
    nb = min(B, 6)
    kk, vv = kv[kvi[:nb].long(), :, :C], kv[kvi[:nb].long(), :, C:]
    qh = q[:nb].float().reshape(nb, T, heads, D).transpose(1, 2)
    kh = kk.float().reshape(nb, Tk, heads, D).transpose(1, 2)
    vh = vv.float().reshape(nb, Tk, heads, D).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(nb, T, C)
    assert torch.isfinite(outs[0]).all()
    assert max_rel(outs[0][:nb], ref) < 2e-3
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("B,T", [(2, 1024), (1, 45), (3, 425), (1, 4096)])
def test_vae_single_head_attention(lib, B, T):
    """one 512-wide head (VAE mid block) on the flash kernel: two CTAs per query tile, any token count"""
    C = 512
    g = torch.Generator(device="cuda").manual_seed(40 + B + T)
    qkv = torch.randn(B, T, 3 * C, device="cuda", generator=g).half()
    out = torch.full((B, T, C), float("nan"), device="cuda", dtype=torch.float16)
    check(lib, lib.dm_op_attention(ptr(qkv[..., :C]), ptr(qkv[..., C:2 * C]), ptr(qkv[..., 2 * C:]), 3 * C, 3 * C, 3 * C,
                                   T * 3 * C, T * 3 * C, T * 3 * C, B, 1, C, T, T, 0, None, ptr(out), C, stream()))
    torch.cuda.synchronize()
    q, k, v = [t.float().reshape(B, T, 1, C).transpose(1, 2) for t in (qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:])]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, C)
    assert torch.isfinite(out).all()
    assert max_rel(out, ref) < 2e-3


def test_attention_large_logits(lib):
    """peaked softmax (|logit| ~ 50): the running-max rescale path in TMEM must stay exact"""
    B, T, D, heads = 1, 512, 40, 8
    C = heads * D
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = (torch.randn(B, T, 3 * C, device="cuda", generator=g) * 4).half()
    out = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
    check(lib, lib.dm_op_attention(ptr(qkv[..., :C]), ptr(qkv[..., C:2 * C]), ptr(qkv[..., 2 * C:]), 3 * C, 3 * C, 3 * C,
                                   T * 3 * C, T * 3 * C, T * 3 * C, B, heads, D, T, T, 0, None, ptr(out), C, stream()))
    q, k, v = [t.float().reshape(B, T, heads, D).transpose(1, 2) for t in (qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:])]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, C)
    assert max_rel(out, ref) < 3e-3


GN_CASES = [(2, 4096, 320, 0, 1, 1e-5), (3, 1024, 640, 320, 1, 1e-5), (2, 256, 1280, 1280, 0, 1e-6), (1, 64, 1280, 640, 1, 1e-5),
            (2, 100, 128, 0, 1, 1e-6), (5, 4096, 640, 320, 1, 1e-5), (1, 1, 320, 0, 1, 1e-5), (700, 64, 1280, 0, 1, 1e-5)]


@pytest.mark.parametrize("case", GN_CASES, ids=lambda c: "-".join(map(str, c)))
def test_groupnorm_silu(lib, case):
    N, HW, C0, C1, silu, eps = case
    C = C0 + C1
    g = torch.Generator(device="cuda").manual_seed(7 + N + HW)
    x = (torch.randn(N, HW, C0, device="cuda", generator=g) * 1.5 + 0.3).half()
    x2 = (torch.randn(N, HW, C1, device="cuda", generator=g) * 0.7 - 0.2).half() if C1 else None
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    out = torch.full((N, HW, C), float("nan"), device="cuda", dtype=torch.float16)
    out2 = torch.empty_like(out)
    check(lib, lib.dm_op_groupnorm(ptr(x), ptr(x2), N, HW, C0, C1, ptr(gamma), ptr(beta), eps, silu, ptr(out), stream()))
    check(lib, lib.dm_op_groupnorm(ptr(x), ptr(x2), N, HW, C0, C1, ptr(gamma), ptr(beta), eps, silu, ptr(out2), stream()))
    xin = x if x2 is None else torch.cat([x, x2], -1)
    ref = F.group_norm(xin.float().permute(0, 2, 1), 32, gamma, beta, eps)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 1)
    assert max_rel(out, ref) < 2e-3
    assert torch.equal(out, out2)  # deterministic reduction: bit-identical run to run


@pytest.mark.parametrize("rows,C", [(4096, 320), (1000, 640), (77, 1280), (1, 320), (40000, 320), (20001, 1280), (300, 512)])
def test_layernorm(lib, rows, C):
    g = torch.Generator(device="cuda").manual_seed(11 + rows)
    x = (torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5).half()
    # the kernel keeps gamma / beta as packed halves: in the engine they ARE fp16 weights (stored as fp32 copies)
    gamma = torch.randn(C, device="cuda", generator=g).half().float()
    beta = torch.randn(C, device="cuda", generator=g).half().float()
    out = torch.full((rows, C), float("nan"), device="cuda", dtype=torch.float16)
    check(lib, lib.dm_op_layernorm(ptr(x), rows, C, ptr(gamma), ptr(beta), 1e-5, ptr(out), stream()))
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    assert max_rel(out, ref) < 2e-3


def test_unsupported_shapes_fail_loudly(lib):
    x = torch.zeros(1, 4, 4, 40, device="cuda", dtype=torch.float16)
    w = torch.zeros(64, 40, device="cuda", dtype=torch.float16)
    out = torch.zeros(16, 64, device="cuda", dtype=torch.float16)
    rc = lib.dm_op_conv(ptr(x), None, 1, 4, 4, 40, 0, ptr(w), 64, 1, 1, 0, None, None, None, ptr(out), 0, 0, 0, 0, stream())
    assert rc != 0 and b"multiples of 64" in lib.dm_last_error()
    q = torch.zeros(1, 16, 8 * 32, device="cuda", dtype=torch.float16)
    rc = lib.dm_op_attention(ptr(q), ptr(q), ptr(q), 256, 256, 256, 4096, 4096, 4096, 1, 8, 32, 16, 16, 0, None, ptr(q), 256, stream())
    assert rc != 0 and b"head_dim" in lib.dm_last_error()
