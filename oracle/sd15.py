"""ORACLE (test infrastructure, not product code) -- a plain-PyTorch restatement of the arithmetic the reference
reaches through `diffusers==0.24.0` on its typicality / DIFT hot path.

PARITY UNPINNED: the reference (ysig/diff-mining @ 0e4c8635) ships no tests, golden vectors or fixtures for
this path, and `diffusers` / `xformers` are not installable in this environment, so this
restatement cannot be checked against outputs of the reference itself.  What pins it instead:
  * the in-repo witnesses of the diffusers control flow, cited per function below
    (paths relative to /root/reference/);
  * structural known-answers: the SD-1.5 U-Net has exactly 859,520,964 parameters in 686 tensors and the VAE
    encoder + quant_conv 34,163,664, with the diffusers state-dict key schema (tests/test_oracle.py);
  * the public SD-1.5 `unet/config.json`, `vae/config.json`, `scheduler/scheduler_config.json` values
    (SURVEY.md appendix A).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Everything is functional: parameters live in a flat dict keyed like the diffusers state dict, so the same dict
feeds this oracle (fp32, "gold") and the CUDA engine (fp16).  `autocast=True` runs the same graph under
torch.autocast(float16) -- the numerics policy of the reference (compute.py:97) -- and is the
"reference-precision" oracle used to interpret the fp16 tolerance.
"""
from __future__ import annotations

import contextlib
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------- configuration
UNET_BLOCK_OUT = (320, 640, 1280, 1280)
UNET_LAYERS_PER_BLOCK = 2
UNET_HEADS = 8  # config "attention_head_dim": 8 is (legacy) the number of heads
UNET_CTX_DIM = 768
UNET_GROUPS = 32
UNET_TIME_DIM = 1280
UNET_DOWN_HAS_ATTN = (True, True, True, False)
UNET_UP_HAS_ATTN = (False, True, True, True)
VAE_BLOCK_OUT = (128, 256, 512, 512)
VAE_SCALING = 0.18215
N_TRAIN_TIMESTEPS = 1000
UNET_PARAM_COUNT = 859_520_964
VAE_ENC_PARAM_COUNT = 34_163_592 + 72


# ----------------------------------------------------------------------------------------------- key schema
def _conv(d, k, cout, cin, ks):
    d[k + ".weight"] = (cout, cin, ks, ks)
    d[k + ".bias"] = (cout,)


def _lin(d, k, cout, cin, bias=True):
    d[k + ".weight"] = (cout, cin)
    if bias:
        d[k + ".bias"] = (cout,)


def _norm(d, k, c):
    d[k + ".weight"] = (c,)
    d[k + ".bias"] = (c,)


def _resnet(d, k, cin, cout, temb=True):
    _norm(d, k + ".norm1", cin)
    _conv(d, k + ".conv1", cout, cin, 3)
    if temb:
        _lin(d, k + ".time_emb_proj", cout, UNET_TIME_DIM)
    _norm(d, k + ".norm2", cout)
    _conv(d, k + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(d, k + ".conv_shortcut", cout, cin, 1)


def _transformer(d, k, c):
    _norm(d, k + ".norm", c)
    _conv(d, k + ".proj_in", c, c, 1)
    t = k + ".transformer_blocks.0"
    _norm(d, t + ".norm1", c)
    for a, kvdim in (("attn1", c), ("attn2", UNET_CTX_DIM)):
        _lin(d, f"{t}.{a}.to_q", c, c, bias=False)
        _lin(d, f"{t}.{a}.to_k", c, kvdim, bias=False)
        _lin(d, f"{t}.{a}.to_v", c, kvdim, bias=False)
        _lin(d, f"{t}.{a}.to_out.0", c, c)
        if a == "attn1":
            _norm(d, t + ".norm2", c)
    _norm(d, t + ".norm3", c)
    _lin(d, t + ".ff.net.0.proj", 8 * c, c)
    _lin(d, t + ".ff.net.2", c, 4 * c)
    _conv(d, k + ".proj_out", c, c, 1)


def unet_param_shapes() -> "OrderedDict[str, Tuple[int, ...]]":
    """diffusers UNet2DConditionModel state-dict schema for SD-1.5."""
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _conv(d, "conv_in", 320, 4, 3)
    _lin(d, "time_embedding.linear_1", UNET_TIME_DIM, 320)
    _lin(d, "time_embedding.linear_2", UNET_TIME_DIM, UNET_TIME_DIM)
    cin = 320
    for i, cout in enumerate(UNET_BLOCK_OUT):
        for j in range(UNET_LAYERS_PER_BLOCK):
            _resnet(d, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
            if UNET_DOWN_HAS_ATTN[i]:
                _transformer(d, f"down_blocks.{i}.attentions.{j}", cout)
        if i < 3:
            _conv(d, f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        cin = cout
    _resnet(d, "mid_block.resnets.0", 1280, 1280)
    _transformer(d, "mid_block.attentions.0", 1280)
    _resnet(d, "mid_block.resnets.1", 1280, 1280)
    rev = tuple(reversed(UNET_BLOCK_OUT))
    out_c = rev[0]
    for i in range(4):
        prev = out_c
        out_c = rev[i]
        in_c = rev[min(i + 1, 3)]
        for j in range(3):
            skip = in_c if j == 2 else out_c
            rin = prev if j == 0 else out_c
            _resnet(d, f"up_blocks.{i}.resnets.{j}", rin + skip, out_c)
            if UNET_UP_HAS_ATTN[i]:
                _transformer(d, f"up_blocks.{i}.attentions.{j}", out_c)
        if i < 3:
            _conv(d, f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    _norm(d, "conv_norm_out", 320)
    _conv(d, "conv_out", 4, 320, 3)
    return d


def vae_encoder_param_shapes() -> "OrderedDict[str, Tuple[int, ...]]":
    """AutoencoderKL `encoder.*` + `quant_conv.*` (SD-1.5 vae/config.json)."""
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _conv(d, "encoder.conv_in", 128, 3, 3)
    cin = 128
    for i, cout in enumerate(VAE_BLOCK_OUT):
        for j in range(2):
            _resnet(d, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb=False)
        if i < 3:
            _conv(d, f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        cin = cout
    _resnet(d, "encoder.mid_block.resnets.0", 512, 512, temb=False)
    a = "encoder.mid_block.attentions.0"
    _norm(d, a + ".group_norm", 512)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(d, f"{a}.{n}", 512, 512)
    _resnet(d, "encoder.mid_block.resnets.1", 512, 512, temb=False)
    _norm(d, "encoder.conv_norm_out", 512)
    _conv(d, "encoder.conv_out", 8, 512, 3)
    _conv(d, "quant_conv", 8, 8, 1)
    return d


def make_synthetic_weights(shapes, seed: int = 0, device="cpu", gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights (no checkpoints are reachable offline).  Fan-in scaled so activations stay O(1)
    through ~60 normalised layers; every value is rounded to fp16 so the fp32 oracle and the fp16 engine consume
    bit-identical parameters."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if k.endswith(".weight") and len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            w = torch.randn(shp, generator=g) * (gain / math.sqrt(fan_in))
        elif k.endswith(".weight"):  # norm gamma
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif ".norm" in k or "group_norm" in k:  # norm beta
            w = 0.1 * torch.randn(shp, generator=g)
        else:  # conv / linear bias
            w = 0.05 * torch.randn(shp, generator=g)
        out[k] = w.half().float().to(device)
    return out


# ----------------------------------------------------------------------------------------------- scheduler
def alphas_cumprod() -> torch.Tensor:
    """scaled_linear betas of SD-1.5's scheduler_config; identical for PNDM/DDPM/DDIM."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, N_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """scheduler.add_noise as called at diffmining/typicality/compute.py:99 and dift.py:190."""
    acp = alphas_cumprod().to(device=x0.device, dtype=x0.dtype)
    a = acp[t] ** 0.5
    b = (1 - acp[t]) ** 0.5
    while a.dim() < x0.dim():
        a = a.unsqueeze(-1)
        b = b.unsqueeze(-1)
    return a * x0 + b * noise


def schedule_tables() -> Tuple[torch.Tensor, torch.Tensor]:
    acp = alphas_cumprod()
    return (acp ** 0.5).contiguous(), ((1 - acp) ** 0.5).contiguous()


# ----------------------------------------------------------------------------------------------- building blocks
class _P:
    """parameter accessor with a key prefix"""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __call__(self, k):
        return self.sd[self.prefix + k]

    def has(self, k):
        return (self.prefix + k) in self.sd

    def sub(self, k):
        return _P(self.sd, self.prefix + k + ".")


def _gn(p: _P, x, eps):
    return F.group_norm(x, UNET_GROUPS, p("weight"), p("bias"), eps)


def _conv2d(p: _P, x, stride=1, padding=1):
    return F.conv2d(x, p("weight"), p("bias"), stride=stride, padding=padding)


def _linear(p: _P, x):
    return F.linear(x, p("weight"), p("bias") if p.has("bias") else None)


def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0) -- the `time_proj` of dift.py:84."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def resnet_block(p: _P, x, temb, eps):
    """ResnetBlock2D.forward; op order witnessed at applications/parallel-dataset/pnp.py:282-359."""
    h = F.silu(_gn(p.sub("norm1"), x, eps))
    h = _conv2d(p.sub("conv1"), h)
    if temb is not None:
        h = h + _linear(p.sub("time_emb_proj"), F.silu(temb))[:, :, None, None]
    h = F.silu(_gn(p.sub("norm2"), h, eps))
    h = _conv2d(p.sub("conv2"), h)
    if p.has("conv_shortcut.weight"):
        x = _conv2d(p.sub("conv_shortcut"), x, padding=0)
    return x + h


def attention(p: _P, x, ctx, heads):
    """Attention.forward with the xformers processor; body witnessed at pnp.py:422-449 (scale d^-0.5, no mask)."""
    q = _linear(p.sub("to_q"), x)
    src = x if ctx is None else ctx
    k = _linear(p.sub("to_k"), src)
    v = _linear(p.sub("to_v"), src)
    B, T, C = q.shape
    d = C // heads
    q = q.view(B, T, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, T, C)
    return _linear(p.sub("to_out.0"), o)


def transformer_2d(p: _P, x, ctx):
    """Transformer2DModel (conv projections) around one BasicTransformerBlock with GEGLU feed-forward."""
    B, C, H, W = x.shape
    res = x
    h = _gn(p.sub("norm"), x, 1e-6)
    h = _conv2d(p.sub("proj_in"), h, padding=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    t = p.sub("transformer_blocks.0")
    n = F.layer_norm(h, (C,), t("norm1.weight"), t("norm1.bias"), 1e-5)
    h = attention(t.sub("attn1"), n, None, UNET_HEADS) + h
    n = F.layer_norm(h, (C,), t("norm2.weight"), t("norm2.bias"), 1e-5)
    h = attention(t.sub("attn2"), n, ctx, UNET_HEADS) + h
    n = F.layer_norm(h, (C,), t("norm3.weight"), t("norm3.bias"), 1e-5)
    proj = _linear(t.sub("ff.net.0.proj"), n)
    val, gate = proj.chunk(2, dim=-1)
    ff = _linear(t.sub("ff.net.2"), val * F.gelu(gate))
    h = ff + h
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    h = _conv2d(p.sub("proj_out"), h, padding=0)
    return h + res


def _autocast(enabled: bool, device_type: str):
    return torch.autocast(device_type=device_type, dtype=torch.float16) if enabled else contextlib.nullcontext()


# ----------------------------------------------------------------------------------------------- U-Net
def unet_forward(sd: Dict[str, torch.Tensor], sample, timesteps, ctx, *, up_ft_index: Optional[int] = None,
                 autocast: bool = False, taps: Optional[dict] = None, prefix: str = ""):
    """UNet2DConditionModel.forward for SD-1.5; top-level control flow follows the in-repo copy at
    diffmining/typicality/dift.py:44-169 (time embedding :84-91, conv_in :104, down loop :107-120, mid :123-130,
    up loop with skip pop and forwarded upsample size :133-165).  With `up_ft_index` set, returns the activation
    after up_blocks[up_ft_index] (including its upsampler) like MyUNet2DConditionModel (dift.py:136-165);
    otherwise returns the epsilon prediction.  `taps`, if given, collects named intermediates."""
    p = _P(sd, prefix)
    eps = 1e-5

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().float()

    with _autocast(autocast, sample.device.type):
        forward_upsample_size = any(s % 8 != 0 for s in sample.shape[-2:])
        t = timesteps
        if t.dim() == 0:
            t = t[None]
        t = t.expand(sample.shape[0])
        t_emb = timestep_embedding(t).to(sd[prefix + "conv_in.weight"].dtype)
        emb = _linear(p.sub("time_embedding.linear_1"), t_emb)
        emb = _linear(p.sub("time_embedding.linear_2"), F.silu(emb))
        h = _conv2d(p.sub("conv_in"), sample)
        tap("conv_in", h)
        skips: List[torch.Tensor] = [h]
        for i in range(4):
            for j in range(UNET_LAYERS_PER_BLOCK):
                h = resnet_block(p.sub(f"down_blocks.{i}.resnets.{j}"), h, emb, eps)
                tap(f"down_blocks.{i}.resnets.{j}", h)
                if UNET_DOWN_HAS_ATTN[i]:
                    h = transformer_2d(p.sub(f"down_blocks.{i}.attentions.{j}"), h, ctx)
                    tap(f"down_blocks.{i}.attentions.{j}", h)
                skips.append(h)
            if i < 3:
                h = _conv2d(p.sub(f"down_blocks.{i}.downsamplers.0.conv"), h, stride=2, padding=1)
                tap(f"down_blocks.{i}.downsamplers.0", h)
                skips.append(h)
        h = resnet_block(p.sub("mid_block.resnets.0"), h, emb, eps)
        h = transformer_2d(p.sub("mid_block.attentions.0"), h, ctx)
        h = resnet_block(p.sub("mid_block.resnets.1"), h, emb, eps)
        tap("mid_block", h)
        for i in range(4):
            if up_ft_index is not None and i > up_ft_index:
                break
            res = skips[-3:]
            skips = skips[:-3]
            upsample_size = skips[-1].shape[2:] if (i < 3 and forward_upsample_size) else None
            for j in range(3):
                h = torch.cat([h, res.pop()], dim=1)
                h = resnet_block(p.sub(f"up_blocks.{i}.resnets.{j}"), h, emb, eps)
                tap(f"up_blocks.{i}.resnets.{j}", h)
                if UNET_UP_HAS_ATTN[i]:
                    h = transformer_2d(p.sub(f"up_blocks.{i}.attentions.{j}"), h, ctx)
                    tap(f"up_blocks.{i}.attentions.{j}", h)
            if i < 3:
                if upsample_size is None:
                    h = F.interpolate(h, scale_factor=2.0, mode="nearest")
                else:
                    h = F.interpolate(h, size=tuple(upsample_size), mode="nearest")
                h = _conv2d(p.sub(f"up_blocks.{i}.upsamplers.0.conv"), h)
                tap(f"up_blocks.{i}.upsamplers.0", h)
            if up_ft_index is not None and i == up_ft_index:
                return h
        h = F.silu(_gn(p.sub("conv_norm_out"), h, eps))
        h = _conv2d(p.sub("conv_out"), h)
        tap("conv_out", h)
    return h


# ----------------------------------------------------------------------------------------------- VAE encoder
def vae_encode_moments(sd: Dict[str, torch.Tensor], x, *, autocast: bool = False, prefix: str = "",
                       taps: Optional[dict] = None):
    """AutoencoderKL.encode up to the posterior moments: Encoder (conv_in, 4 DownEncoderBlock2D, mid block with a
    single-head 512-d attention, GN+SiLU, conv_out) then quant_conv.  Reached from compute.py:93 / dift.py:187.
    Returns (mean, logvar) with logvar clamped to [-30, 20] (DiagonalGaussianDistribution)."""
    p = _P(sd, prefix)
    eps = 1e-6

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().float()

    with _autocast(autocast, x.device.type):
        h = _conv2d(p.sub("encoder.conv_in"), x)
        tap("encoder.conv_in", h)
        for i in range(4):
            for j in range(2):
                h = resnet_block(p.sub(f"encoder.down_blocks.{i}.resnets.{j}"), h, None, eps)
            tap(f"encoder.down_blocks.{i}", h)
            if i < 3:
                h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
                h = _conv2d(p.sub(f"encoder.down_blocks.{i}.downsamplers.0.conv"), h, stride=2, padding=0)
                tap(f"encoder.down_blocks.{i}.downsamplers.0", h)
        h = resnet_block(p.sub("encoder.mid_block.resnets.0"), h, None, eps)
        a = p.sub("encoder.mid_block.attentions.0")
        B, C, H, W = h.shape
        res = h
        n = _gn(a.sub("group_norm"), h.view(B, C, H * W), eps).transpose(1, 2)
        o = attention(a, n, None, 1)
        h = o.transpose(1, 2).reshape(B, C, H, W) + res
        tap("encoder.mid_block.attentions.0", h)
        h = resnet_block(p.sub("encoder.mid_block.resnets.1"), h, None, eps)
        h = F.silu(_gn(p.sub("encoder.conv_norm_out"), h, eps))
        h = _conv2d(p.sub("encoder.conv_out"), h)
        moments = _conv2d(p.sub("quant_conv"), h, padding=0)
        mean, logvar = torch.chunk(moments, 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean, logvar


def vae_sample(mean, logvar, eps_draw, scaling: float = VAE_SCALING):
    """latent_dist.sample() * scaling_factor (compute.py:93): mean + exp(0.5*logvar) * eps, in fp32 (the exp is
    promoted to fp32 by autocast)."""
    std = torch.exp(0.5 * logvar.float())
    return (mean.float() + std * eps_draw.float()) * scaling


def count_params(shapes) -> int:
    n = 0
    for s in shapes.values():
        k = 1
        for v in s:
            k *= v
        n += k
    return n
