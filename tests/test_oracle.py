"""CPU tests of the oracle (the parity checker itself): structural known-answers, committed golden vectors,
and the small pieces of arithmetic whose definition is public (scheduler, sinusoid, T reduction)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import sd15

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


def test_unet_param_count_and_schema():
    shapes = sd15.unet_param_shapes()
    assert len(shapes) == 686
    assert sd15.count_params(shapes) == 859_520_964  # SD-1.5 UNet2DConditionModel
    for k in ("time_embedding.linear_1.weight", "conv_in.weight", "down_blocks.0.resnets.0.time_emb_proj.weight",
              "down_blocks.1.resnets.0.conv_shortcut.weight", "down_blocks.2.downsamplers.0.conv.weight",
              "down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight",
              "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj.weight", "up_blocks.1.upsamplers.0.conv.weight",
              "up_blocks.3.attentions.2.proj_out.bias", "conv_norm_out.weight", "conv_out.bias"):
        assert k in shapes, k
    assert "down_blocks.3.attentions.0.norm.weight" not in shapes       # DownBlock2D has no attention
    assert "down_blocks.3.downsamplers.0.conv.weight" not in shapes     # final down block does not downsample
    assert "up_blocks.0.attentions.0.norm.weight" not in shapes         # UpBlock2D
    assert shapes["up_blocks.1.resnets.2.conv1.weight"] == (1280, 1920, 3, 3)
    assert shapes["up_blocks.3.resnets.0.conv1.weight"] == (320, 960, 3, 3)
    assert shapes["down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight"] == (320, 320)
    assert "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.bias" not in shapes
    assert shapes["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"] == (320, 768)


def test_vae_param_count_and_schema():
    shapes = sd15.vae_encoder_param_shapes()
    assert sd15.count_params(shapes) == 34_163_592 + 72
    assert shapes["encoder.conv_out.weight"] == (8, 512, 3, 3)
    assert shapes["quant_conv.weight"] == (8, 8, 1, 1)
    assert shapes["encoder.mid_block.attentions.0.to_q.bias"] == (512,)
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in shapes


def test_golden_vectors(unet_weights, vae_weights):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLD), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(GOLD)
    x, t, ctx, img, x_odd = mg.inputs()
    np.testing.assert_array_equal(gold["w_probe"][0], unet_weights["conv_in.weight"].flatten()[:8].numpy())
    with torch.no_grad():
        eps = sd15.unet_forward(unet_weights, x, t, ctx)
        eps_odd = sd15.unet_forward(unet_weights, x_odd, t[:1], ctx[:1])
        mean, logvar = sd15.vae_encode_moments(vae_weights, img)
    # same code, same seeds, fp32 CPU: only thread-count dependent summation order may differ
    np.testing.assert_allclose(eps.numpy(), gold["eps"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(eps_odd.numpy(), gold["eps_odd"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(mean.numpy(), gold["vae_mean"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(logvar.numpy(), gold["vae_logvar"], rtol=2e-4, atol=2e-5)


def test_dift_early_exit_shapes(unet_weights):
    x = torch.randn(1, 4, 16, 16)
    t = torch.tensor([161])
    ctx = torch.randn(1, 77, 768)
    with torch.no_grad():
        for idx, (c, s) in {0: (1280, 4), 1: (1280, 8), 2: (640, 16)}.items():
            f = sd15.unet_forward(unet_weights, x, t, ctx, up_ft_index=idx)
            assert f.shape == (1, c, s, s)


def test_scheduler_tables():
    acp = sd15.alphas_cumprod()
    assert acp.shape == (1000,)
    assert abs(acp[0].item() - (1 - 0.00085)) < 1e-6
    assert abs(acp[-1].item() - 0.0047) < 2e-4  # SD-1.5's terminal alpha_bar
    assert torch.all(acp[1:] < acp[:-1])
    x0, eps = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8)
    t = torch.tensor([0, 500, 999])
    noisy = sd15.add_noise(x0, eps, t)
    a, b = sd15.schedule_tables()
    ref = a[t].view(3, 1, 1, 1) * x0 + b[t].view(3, 1, 1, 1) * eps
    torch.testing.assert_close(noisy, ref)
    # a^2 + b^2 = 1
    torch.testing.assert_close(a * a + b * b, torch.ones(1000), atol=1e-6, rtol=0)


def test_timestep_embedding():
    t = torch.tensor([0, 1, 999])
    e = sd15.timestep_embedding(t)
    assert e.shape == (3, 320)
    torch.testing.assert_close(e[0, :160], torch.ones(160))   # cos(0) first: flip_sin_to_cos
    torch.testing.assert_close(e[0, 160:], torch.zeros(160))
    assert abs(e[1, 160].item() - math.sin(1.0)) < 1e-6
    assert abs(e[2, 159].item() - math.cos(999 * math.exp(-math.log(10000) * 159 / 160))) < 1e-4


def test_vae_sample_and_sizes(vae_weights):
    img = torch.rand(1, 3, 40, 72) * 2 - 1
    with torch.no_grad():
        m, lv = sd15.vae_encode_moments(vae_weights, img)
    assert m.shape == (1, 4, 5, 9) and lv.shape == (1, 4, 5, 9)
    assert lv.min() >= -30 and lv.max() <= 20
    z0 = sd15.vae_sample(m, lv, torch.zeros_like(m))
    torch.testing.assert_close(z0, m * 0.18215)


def test_synthetic_weights_are_fp16_representable(unet_weights):
    for k in ("conv_in.weight", "mid_block.resnets.0.conv1.weight", "conv_out.bias"):
        w = unet_weights[k]
        assert torch.equal(w, w.half().float())


def test_consumer_oracle_matches_pandas_restatement():
    """oracle/consumers.py against a literal pandas transcription of df_D + get_non_overlapping (cluster.py:193-204,
    utils.py:83-102) on a small random loss grid"""
    import pandas as pd

    from oracle import consumers

    rng = np.random.default_rng(0)
    grid = rng.random((3, 2, 4, 6, 7)).astype(np.float16)
    H, W, kx, ky, k = 48, 56, 8, 8, 5
    D = consumers.patch_scores(grid, H, W, kx, ky)
    assert D.shape == (H - kx + 1, W - ky + 1)
    df = pd.DataFrame([(i, j, i + kx, j + ky, D[i, j]) for i in range(D.shape[0]) for j in range(D.shape[1])],
                      columns=["x_start", "y_start", "x_end", "y_end", "D"])
    df = df.sort_values(by=["D"], ascending=False, kind="stable").reset_index(drop=True)
    picked = []
    while len(picked) < k:
        picked.append(df.iloc[0])
        last = picked[-1]
        df = df[~((df["x_start"] <= last["x_end"]) & (df["x_end"] >= last["x_start"]) & (df["y_start"] <= last["y_end"])
                  & (df["y_end"] >= last["y_start"]))].reset_index(drop=True)
        if df.shape[0] == 0:
            break
    ours = consumers.non_overlapping_topk(D, kx, ky, k)
    assert [(int(r["x_start"]), int(r["y_start"])) for r in picked] == [(r[0], r[1]) for r in ours]
    # boxes are mutually non-overlapping under the inclusive test
    for a in range(len(ours)):
        for b in range(a):
            assert abs(ours[a][0] - ours[b][0]) > kx or abs(ours[a][1] - ours[b][1]) > ky
