"""Generates tests/golden/reference_golden.npz by EXECUTING THE REFERENCE'S OWN CODE, imported unmodified from
/root/reference in this container (it does not exist on the GPU box; only the vectors travel).

What is reference code and what is not
--------------------------------------
The reference's arithmetic is split between its own files and `diffusers==0.24.0` / `xformers` (not installable here).
This script imports the reference modules with *stub* `diffusers` / `xformers` / `matplotlib` / `umap` / `skimage` packages
that only provide the names the imports need, builds a tree of plain `torch.nn` modules carrying the oracle's seeded
synthetic weights under the diffusers attribute names, and then runs, unmodified:

  * `MyUNet2DConditionModel.forward`                      diffmining/typicality/dift.py:24-169
        (time embedding cast, conv_in, down loop, skip bookkeeping, mid, up loop with skip pop order and forwarded
        upsample size, DIFT early exit)
  * `OneStepSDPipeline.__call__`                           diffmining/typicality/dift.py:172-192
        (vae.encode(...).latent_dist.sample() * scaling_factor -> randn_like -> add_noise -> unet)
  * the ResnetBlock2D forward  `conv_forward`              diffmining/applications/parallel-dataset/pnp.py:277-359
        bound to all 22 U-Net ResNets and the 10 VAE-encoder ResNets
  * the Attention forward      `sa_forward`                pnp.py:378-459
        bound to all 32 U-Net attentions (self and cross) and the VAE mid-block attention; its
        `xformers.ops.memory_efficient_attention` call is served by a stub computing softmax(q k^T * scale) v in fp32
  * `SD.compute_loss`, `D.noising`, `D.compute_losses`     diffmining/typicality/compute.py:95-160
  * `Cluster.load_typicality_norm` / `load_typicality`     diffmining/typicality/cluster.py:112-137
  * `pool`, `sort`, `get_non_overlapping`                  diffmining/typicality/utils.py:74-102

What the stubs restate (diffusers-internal glue that the reference never spells out): the block containers
(CrossAttnDownBlock2D, UNetMidBlock2DCrossAttn, UpBlock2D, ...), Transformer2DModel / BasicTransformerBlock / GEGLU,
Downsample2D / Upsample2D, Timesteps / TimestepEmbedding, head_to_batch_dim / batch_to_head_dim, the VAE Encoder's
top-level sequence, DiagonalGaussianDistribution and scheduler.add_noise.  Those follow the published diffusers 0.24
behaviour and SD-1.5 configs exactly as oracle/sd15.py does.

The oracle is then PINNED by tests/test_oracle_pinned.py: oracle/sd15.py and oracle/consumers.py must reproduce these
reference-executed vectors to fp32 rounding.  Run from the repo root:  python tests/golden/make_reference_golden.py"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------ import stubs
class _ModelBase(nn.Module):
    """stands in for diffusers' ModelMixin: only `.dtype` is used by the reference forward (dift.py:89)"""

    @property
    def dtype(self):
        return next(self.parameters()).dtype


def _mem_eff_attention(query, key, value, attn_bias=None, op=None, scale=None):
    """xformers.ops.memory_efficient_attention for [B*heads, T, d] inputs: softmax(q k^T * scale) v (published semantics)"""
    assert attn_bias is None
    scale = query.shape[-1] ** -0.5 if scale is None else scale
    s = torch.baddbmm(torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype), query, key.transpose(1, 2),
                      beta=0, alpha=scale)
    return torch.bmm(s.softmax(dim=-1), value)


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    blank = lambda n: type(n, (object,), {})  # noqa: E731
    mod("diffusers", StableDiffusionPipeline=blank("StableDiffusionPipeline"), DDIMScheduler=blank("DDIMScheduler"),
        AutoencoderKL=blank("AutoencoderKL"), UNet2DConditionModel=_ModelBase)
    mod("diffusers.models")
    mod("diffusers.models.unet_2d_condition", UNet2DConditionModel=_ModelBase)
    mod("diffusers.models.attention_processor", Attention=blank("Attention"))
    mod("diffusers.utils", USE_PEFT_BACKEND=True)
    x = mod("xformers")
    x.ops = mod("xformers.ops", memory_efficient_attention=_mem_eff_attention)
    mp = mod("matplotlib")
    mp.pyplot = mod("matplotlib.pyplot")
    mod("umap")
    sk = mod("skimage")
    sk.exposure = mod("skimage.exposure")
    sk.filters = mod("skimage.filters")


def load_ref(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


# ------------------------------------------------------------------------------------------------ stub module tree
class Resnet(nn.Module):
    """attribute surface of diffusers' ResnetBlock2D as the reference's conv_forward reads it (pnp.py:277-359)"""

    def __init__(self, cin, cout, temb, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(1280, cout) if temb else None
        self.norm2 = nn.GroupNorm(32, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None
        self.nonlinearity = F.silu
        self.upsample = self.downsample = None
        self.time_embedding_norm = "default"
        self.skip_time_act = False
        self.output_scale_factor = 1.0


class Attention(nn.Module):
    """attribute surface of diffusers' Attention as the reference's sa_forward reads it (pnp.py:378-459)"""

    def __init__(self, C, kv_dim, heads, bias, out_bias=True, group_norm_eps=None, residual=False):
        super().__init__()
        self.heads = heads
        self.scale = (C // heads) ** -0.5
        self.to_q = nn.Linear(C, C, bias=bias)
        self.to_k = nn.Linear(kv_dim, C, bias=bias)
        self.to_v = nn.Linear(kv_dim, C, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(C, C, bias=out_bias), nn.Dropout(0.0)])
        self.group_norm = nn.GroupNorm(32, C, eps=group_norm_eps) if group_norm_eps else None
        self.spatial_norm = None
        self.norm_cross = False
        self.residual_connection = residual
        self.rescale_output_factor = 1.0
        self.processor = types.SimpleNamespace(attention_op=None)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        assert attention_mask is None
        return None

    def head_to_batch_dim(self, t):  # diffusers Attention.head_to_batch_dim (out_dim=3)
        B, T, C = t.shape
        return t.reshape(B, T, self.heads, C // self.heads).permute(0, 2, 1, 3).reshape(B * self.heads, T, C // self.heads)

    def batch_to_head_dim(self, t):
        BH, T, d = t.shape
        return t.reshape(BH // self.heads, self.heads, T, d).permute(0, 2, 1, 3).reshape(BH // self.heads, T, d * self.heads)


class GEGLU(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.proj = nn.Linear(C, 8 * C)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class TransformerBlock(nn.Module):  # BasicTransformerBlock (restated glue)
    def __init__(self, C):
        super().__init__()
        self.norm1 = nn.LayerNorm(C)
        self.attn1 = Attention(C, C, 8, bias=False)
        self.norm2 = nn.LayerNorm(C)
        self.attn2 = Attention(C, 768, 8, bias=False)
        self.norm3 = nn.LayerNorm(C)
        self.ff = nn.Module()
        self.ff.net = nn.ModuleList([GEGLU(C), nn.Dropout(0.0), nn.Linear(4 * C, C)])

    def forward(self, h, ctx):
        h = self.attn1.forward(self.norm1(h), encoder_hidden_states=None) + h
        h = self.attn2.forward(self.norm2(h), encoder_hidden_states=ctx) + h
        n = self.norm3(h)
        for m in self.ff.net:
            n = m(n)
        return n + h


class Transformer2D(nn.Module):  # Transformer2DModel, use_linear_projection=False (restated glue)
    def __init__(self, C):
        super().__init__()
        self.norm = nn.GroupNorm(32, C, eps=1e-6)
        self.proj_in = nn.Conv2d(C, C, 1)
        self.transformer_blocks = nn.ModuleList([TransformerBlock(C)])
        self.proj_out = nn.Conv2d(C, C, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.proj_in(self.norm(x))
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.transformer_blocks[0](h, ctx)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(h) + x


class Downsample(nn.Module):
    def __init__(self, C, pad):
        super().__init__()
        self.conv = nn.Conv2d(C, C, 3, stride=2, padding=pad)
        self.pad = pad

    def forward(self, x):
        if self.pad == 0:  # VAE encoder: F.pad (0,1,0,1) then pad-0 conv
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.conv = nn.Conv2d(C, C, 3, padding=1)

    def forward(self, x, output_size=None):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest") if output_size is None else F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, attn, down):
        super().__init__()
        self.has_cross_attention = attn
        self.resnets = nn.ModuleList([Resnet(cin if j == 0 else cout, cout, True, 1e-5) for j in range(2)])
        if attn:
            self.attentions = nn.ModuleList([Transformer2D(cout) for _ in range(2)])
        self.downsamplers = nn.ModuleList([Downsample(cout, 1)]) if down else None

    def forward(self, hidden_states, temb, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        out = ()
        for j, r in enumerate(self.resnets):
            hidden_states = r.forward(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[j](hidden_states, encoder_hidden_states)
            out += (hidden_states,)
        if self.downsamplers is not None:
            hidden_states = self.downsamplers[0](hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class MidBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(1280, 1280, True, 1e-5) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2D(1280)])

    def forward(self, hidden_states, temb, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
        h = self.resnets[0].forward(hidden_states, temb)
        h = self.attentions[0](h, encoder_hidden_states)
        return self.resnets[1].forward(h, temb)


class UpBlock(nn.Module):
    def __init__(self, cins, cout, attn, up):
        super().__init__()
        self.has_cross_attention = attn
        self.resnets = nn.ModuleList([Resnet(c, cout, True, 1e-5) for c in cins])
        if attn:
            self.attentions = nn.ModuleList([Transformer2D(cout) for _ in range(3)])
        self.upsamplers = nn.ModuleList([Upsample(cout)]) if up else None

    def forward(self, hidden_states, temb, res_hidden_states_tuple, encoder_hidden_states=None, cross_attention_kwargs=None,
                upsample_size=None, attention_mask=None):
        for j, r in enumerate(self.resnets):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = r.forward(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[j](hidden_states, encoder_hidden_states)
        if self.upsamplers is not None:
            hidden_states = self.upsamplers[0](hidden_states, upsample_size)
        return hidden_states


class TimeProj(nn.Module):  # Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0)
    def forward(self, t):
        half = 160
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32) / half
        emb = t[:, None].float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimeEmbedding(nn.Module):
    def __init__(self):
        super().__init__()
        self.linear_1 = nn.Linear(320, 1280)
        self.linear_2 = nn.Linear(1280, 1280)

    def forward(self, sample, condition=None):
        return self.linear_2(F.silu(self.linear_1(sample)))


def build_unet(my_unet_cls, usd):
    u = my_unet_cls.__new__(my_unet_cls)
    nn.Module.__init__(u)
    u.config = types.SimpleNamespace(center_input_sample=False, class_embed_type=None)
    u.num_upsamplers = 3
    u.time_proj = TimeProj()
    u.time_embedding = TimeEmbedding()
    u.class_embedding = None
    u.conv_in = nn.Conv2d(4, 320, 3, padding=1)
    ch = (320, 640, 1280, 1280)
    u.down_blocks = nn.ModuleList([DownBlock(ch[max(i - 1, 0)], ch[i], i < 3, i < 3) for i in range(4)])
    u.mid_block = MidBlock()
    rev = (1280, 1280, 640, 320)
    ups, out_c = [], rev[0]
    for i in range(4):
        prev, out_c, in_c = out_c, rev[i], rev[min(i + 1, 3)]
        cins = [(prev if j == 0 else out_c) + (in_c if j == 2 else out_c) for j in range(3)]
        ups.append(UpBlock(cins, out_c, i > 0, i < 3))
    u.up_blocks = nn.ModuleList(ups)
    u.conv_norm_out = nn.GroupNorm(32, 320, eps=1e-5)
    u.conv_out = nn.Conv2d(320, 4, 3, padding=1)
    missing, unexpected = u.load_state_dict(usd, strict=True), None
    del missing, unexpected
    return u.eval()


class VaeEncoder(nn.Module):  # diffusers Encoder + quant_conv, top-level sequence restated
    def __init__(self):
        super().__init__()
        e = nn.Module()
        e.conv_in = nn.Conv2d(3, 128, 3, padding=1)
        ch = (128, 256, 512, 512)
        blocks = []
        for i in range(4):
            b = nn.Module()
            b.resnets = nn.ModuleList([Resnet(ch[max(i - 1, 0)] if j == 0 else ch[i], ch[i], False, 1e-6) for j in range(2)])
            if i < 3:
                b.downsamplers = nn.ModuleList([Downsample(ch[i], 0)])
            blocks.append(b)
        e.down_blocks = nn.ModuleList(blocks)
        e.mid_block = nn.Module()
        e.mid_block.resnets = nn.ModuleList([Resnet(512, 512, False, 1e-6) for _ in range(2)])
        e.mid_block.attentions = nn.ModuleList([Attention(512, 512, 1, bias=True, group_norm_eps=1e-6, residual=True)])
        e.conv_norm_out = nn.GroupNorm(32, 512, eps=1e-6)
        e.conv_out = nn.Conv2d(512, 8, 3, padding=1)
        self.encoder = e
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.config = types.SimpleNamespace(scaling_factor=0.18215)

    def moments(self, x):
        e = self.encoder
        h = e.conv_in(x)
        for i, b in enumerate(e.down_blocks):
            for r in b.resnets:
                h = r.forward(h, None)
            if i < 3:
                h = b.downsamplers[0](h)
        h = e.mid_block.resnets[0].forward(h, None)
        h = e.mid_block.attentions[0].forward(h)
        h = e.mid_block.resnets[1].forward(h, None)
        h = e.conv_out(F.silu(e.conv_norm_out(h)))
        mean, logvar = torch.chunk(self.quant_conv(h), 2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)

    def encode(self, x):  # AutoencoderKL.encode(x).latent_dist.sample(): DiagonalGaussianDistribution (restated)
        mean, logvar = self.moments(x)
        dist = types.SimpleNamespace(sample=lambda: mean + torch.exp(0.5 * logvar) * torch.randn(mean.shape, dtype=mean.dtype))
        return types.SimpleNamespace(latent_dist=dist)


class Scheduler:
    """scheduler.add_noise (DDPM/DDIM/PNDM share it) for SD-1.5's scheduler_config (restated)"""
    num_train_timesteps = 1000

    def __init__(self):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(dtype=original_samples.dtype)
        a = (acp[timesteps] ** 0.5).flatten()
        b = ((1 - acp[timesteps]) ** 0.5).flatten()
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise


def bind_reference_forwards(pnp, unet, vae):
    """hand every ResNet / Attention of the stub trees to the reference's register_* functions so that their `.forward`
    becomes the reference's own closure (pnp.py:363-373, 463-476 install them on mid_block / up_blocks entries)"""
    resnets = [m for m in list(unet.modules()) + list(vae.modules()) if isinstance(m, Resnet)]
    attns = [m for m in list(unet.modules()) + list(vae.modules()) if isinstance(m, Attention)]
    fake = types.SimpleNamespace(unet=types.SimpleNamespace(
        mid_block=types.SimpleNamespace(
            resnets=resnets,
            attentions=[types.SimpleNamespace(transformer_blocks=[types.SimpleNamespace(attn1=a)]) for a in attns]),
        up_blocks=[]))
    pnp.register_conv_control_efficient(fake, None, lambda res, block: True)
    pnp.register_attention_control_efficient(fake, None, lambda res, block: True)
    assert len(resnets) == 22 + 10 and len(attns) == 32 + 1
    for m in resnets + attns:
        assert m.forward.__module__ == pnp.__name__ and m.injection_schedule is None
    return len(resnets), len(attns)


# ------------------------------------------------------------------------------------------------ inputs (shared with tests)
def unet_inputs():
    g = torch.Generator().manual_seed(4242)
    x = torch.randn(2, 4, 8, 12, generator=g)
    t = torch.tensor([37, 911])
    ctx = torch.randn(2, 77, 768, generator=g)
    x_odd = torch.randn(1, 4, 9, 13, generator=g)   # not a multiple of 8: forwarded upsample sizes (dift.py:48-57,144-147)
    return x, t, ctx, x_odd


def image_inputs():
    g = torch.Generator().manual_seed(777)
    img = torch.rand(2, 3, 64, 48, generator=g) * 2 - 1
    return img


def pil_image():
    from PIL import Image

    rng = np.random.RandomState(11)
    return Image.fromarray(rng.randint(0, 256, (48, 64, 3), dtype=np.uint8))


def consumer_grid():
    rng = np.random.RandomState(5)
    return (rng.rand(3, 2, 4, 6, 8).astype(np.float32) ** 2).astype(np.float16)  # [N, n_cond, 4, h, w] like compute.py:160


def main():
    from oracle import sd15

    torch.set_num_threads(os.cpu_count() or 1)
    install_stubs()
    dift = load_ref("ref_dift", "diffmining/typicality/dift.py")
    pnp = load_ref("ref_pnp", "diffmining/applications/parallel-dataset/pnp.py")
    comp = load_ref("ref_compute", "diffmining/typicality/compute.py")
    utils = load_ref("diffmining.typicality.utils", "diffmining/typicality/utils.py")

    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    vsd = sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1)
    unet = build_unet(dift.MyUNet2DConditionModel, usd)
    vae = VaeEncoder()
    vae.load_state_dict(vsd, strict=True)
    vae.eval()
    n_res, n_att = bind_reference_forwards(pnp, unet, vae)
    print(f"reference forwards bound: {n_res} ResnetBlock2D, {n_att} Attention")
    out = {}
    x, t, ctx, x_odd = unet_inputs()
    with torch.no_grad():
        # ---- (1) MyUNet2DConditionModel.forward: every DIFT tap, incl. the last block (input of conv_norm_out)
        for idx in (0, 1, 2, 3):
            out[f"up_ft{idx}"] = unet.forward(x, t, [idx], encoder_hidden_states=ctx)["up_ft"][idx].numpy()
        up3 = unet.forward(x, t, [3], encoder_hidden_states=ctx)["up_ft"][3]
        out["eps"] = unet.conv_out(F.silu(unet.conv_norm_out(up3))).numpy()   # post-process of UNet2DConditionModel (restated)
        up3_odd = unet.forward(x_odd, t[:1], [3], encoder_hidden_states=ctx[:1])["up_ft"][3]
        out["up_ft3_odd"] = up3_odd.numpy()
        out["up_ft1_odd"] = unet.forward(x_odd, t[:1], [1], encoder_hidden_states=ctx[:1])["up_ft"][1].numpy()
        out["eps_odd"] = unet.conv_out(F.silu(unet.conv_norm_out(up3_odd))).numpy()
        # scalar python timestep path (dift.py:69-79)
        out["up_ft1_scalar_t"] = unet.forward(x, 261, [1], encoder_hidden_states=ctx)["up_ft"][1].numpy()

        # ---- (2) VAE encoder moments through the reference ResNet / Attention forwards
        img = image_inputs()
        mean, logvar = vae.moments(img)
        out["vae_mean"], out["vae_logvar"] = mean.numpy(), logvar.numpy()

        # ---- (3) OneStepSDPipeline.__call__ (dift.py:172-192)
        pipe = object.__new__(dift.OneStepSDPipeline)
        pipe.vae, pipe.unet, pipe.scheduler = vae, unet, Scheduler()
        type(pipe)._execution_device = property(lambda self: torch.device("cpu"))
        torch.manual_seed(99)
        res = pipe(img_tensor=img[:1].repeat(2, 1, 1, 1), t=161, up_ft_indices=[1], prompt_embeds=ctx)
        out["onestep_up_ft1"] = res["up_ft"][1].numpy()
        out["onestep_mean"] = res["up_ft"][1].mean(0, keepdim=True).numpy()   # SDFeaturizer.forward's ensemble mean (dift.py:231)

        # ---- (4) SD.compute_loss / D.noising / D.compute_losses (compute.py:95-160)
        sd = object.__new__(comp.SD)
        sd.device = torch.device("cpu")
        sd.scheduler = Scheduler()
        sd.vae = vae
        sd.model = types.SimpleNamespace(unet=lambda noisy, ts, c: types.SimpleNamespace(
            sample=unet.conv_out(F.silu(unet.conv_norm_out(unet.forward(noisy, ts, [3], encoder_hidden_states=c)["up_ft"][3])))))
        d = comp.D(sd, "/tmp/unused", "cars", seed=42, N=5, t_min=0.1, t_max=0.7)
        pil = pil_image()
        embeds = torch.stack([ctx[0], ctx[1], ctx[0] * 0.5])     # 3 conditions (X-ray style n_cond > 2)
        torch.manual_seed(7)                                       # the posterior draw of encode_vae is unseeded in the reference
        grid = d.compute_losses(pil, embeds, B=2)                  # ragged micro-batches: 2, 2, 1
        out["losses_grid"] = grid.numpy()
        torch.manual_seed(7)
        out["latent"] = sd.encode_vae(d.load_image(pil)).numpy()
        torch.manual_seed(42)
        nz, ts = zip(*[d.noising(torch.from_numpy(out["latent"])) for _ in range(5)])
        out["draw_noise"], out["draw_t"] = torch.cat(nz).numpy(), torch.cat(ts).numpy()
        out["compute_loss_rows"] = sd.compute_loss(torch.from_numpy(out["latent"]), torch.cat(nz)[:2], torch.cat(ts)[:2],
                                                   ctx[1][None].expand(2, -1, -1)).numpy()

    # ---- (5) consumers: Cluster.load_typicality_norm / load_typicality + utils.pool / sort / get_non_overlapping
    import pandas as pd

    sys.modules["diffmining"] = types.ModuleType("diffmining")
    sys.modules["diffmining.typicality"] = types.ModuleType("diffmining.typicality")
    sys.modules["diffmining.typicality.dift"] = dift
    sys.modules["diffmining.typicality.compute"] = types.SimpleNamespace(Typicality=object)
    cluster = load_ref("ref_cluster", "diffmining/typicality/cluster.py")
    cg = consumer_grid()
    H, W, kx, ky = 48, 64, 16, 16
    fake = types.SimpleNamespace(load_image=lambda p: types.SimpleNamespace(size=(W, H)), device="cpu", kx=kx, ky=ky)
    out["T_norm"] = cluster.Cluster.load_typicality_norm(fake, lambda p: cg, "x.png")
    Dm = cluster.Cluster.load_typicality.__wrapped__(fake, lambda p: cg, "x.png") if hasattr(cluster.Cluster.load_typicality, "__wrapped__") \
        else cluster.Cluster.load_typicality(fake, lambda p: cg, "x.png")
    out["D_map"] = Dm
    rows = [("x.png", i, j, i + kx, j + ky, Dm[i, j], "real") for i in range(Dm.shape[0]) for j in range(Dm.shape[1])]   # cluster.py:193
    df = utils.sort(pd.DataFrame(rows, columns=["seed", "x_start", "y_start", "x_end", "y_end", "D", "origin"]), "D", ascending=False)
    top = utils.get_non_overlapping(df, k_per_image=5)
    top = pd.DataFrame(top) if not isinstance(top, pd.DataFrame) else top
    out["topk_boxes"] = top[["x_start", "y_start", "x_end", "y_end"]].to_numpy().astype(np.int64)
    out["topk_scores"] = top["D"].to_numpy().astype(np.float64)

    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("written", path, {k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
