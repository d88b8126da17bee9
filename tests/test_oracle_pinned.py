"""PINS THE ORACLE: oracle/sd15.py and oracle/consumers.py must reproduce, to fp32 rounding, vectors produced by EXECUTING
the reference's own code (tests/golden/make_reference_golden.py, run in the build container where /root/reference
exists): MyUNet2DConditionModel.forward (dift.py:24-169), OneStepSDPipeline.__call__ (dift.py:172-192), the
ResnetBlock2D / Attention forwards of pnp.py:277-459 on every ResNet and attention of the U-Net and VAE encoder,
SD.compute_loss / D.noising / D.compute_losses (compute.py:95-160), Cluster.load_typicality(_norm) (cluster.py:112-137)
and utils.pool / sort / get_non_overlapping (utils.py:74-102).  CPU only, a few seconds."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import consumers, sd15

GDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GDIR, "reference_golden.npz"))


@pytest.fixture(scope="module")
def gen():
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GDIR, "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)   # defines the seeded inputs; touches neither /root/reference nor the stubs
    return m


def close(a, b, what, rtol=2e-4, atol=2e-5):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, f"{what}: {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    assert torch.allclose(a, b, rtol=rtol, atol=atol * max(scale, 1.0)), f"{what}: max abs err {err:.3e} (scale {scale:.3e})"


def test_unet_forward_matches_reference_execution(ref, gen, unet_weights):
    x, t, ctx, x_odd = gen.unet_inputs()
    with torch.no_grad():
        for idx in (0, 1, 2):
            close(sd15.unet_forward(unet_weights, x, t, ctx, up_ft_index=idx), ref[f"up_ft{idx}"], f"up_ft[{idx}]")
        taps = {}
        eps = sd15.unet_forward(unet_weights, x, t, ctx, taps=taps)
        close(taps["up_blocks.3.attentions.2"], ref["up_ft3"], "up_ft[3]")
        close(eps, ref["eps"], "eps")
        # latent not a multiple of 8: forwarded upsample sizes (dift.py:48-57,144-147)
        taps = {}
        eps_odd = sd15.unet_forward(unet_weights, x_odd, t[:1], ctx[:1], taps=taps)
        close(taps["up_blocks.3.attentions.2"], ref["up_ft3_odd"], "up_ft[3] odd")
        close(eps_odd, ref["eps_odd"], "eps odd")
        close(sd15.unet_forward(unet_weights, x_odd, t[:1], ctx[:1], up_ft_index=1), ref["up_ft1_odd"], "up_ft[1] odd")
        # python-int timestep broadcast (dift.py:69-82)
        close(sd15.unet_forward(unet_weights, x, torch.tensor(261), ctx, up_ft_index=1), ref["up_ft1_scalar_t"], "scalar t")


def test_vae_and_onestep_pipeline_match_reference_execution(ref, gen, unet_weights, vae_weights):
    img = gen.image_inputs()
    _, _, ctx, _ = gen.unet_inputs()
    with torch.no_grad():
        mean, logvar = sd15.vae_encode_moments(vae_weights, img)
        close(mean, ref["vae_mean"], "vae mean")
        close(logvar, ref["vae_logvar"], "vae logvar")
        # OneStepSDPipeline.__call__: same RNG stream -- posterior draw, then randn_like(latents) (dift.py:187-189)
        torch.manual_seed(99)
        m2, lv2 = sd15.vae_encode_moments(vae_weights, img[:1].repeat(2, 1, 1, 1))
        lat = sd15.vae_sample(m2, lv2, torch.randn(m2.shape))
        noise = torch.randn_like(lat)
        tt = torch.full((2,), 161)
        ft = sd15.unet_forward(unet_weights, sd15.add_noise(lat, noise, tt), tt, ctx, up_ft_index=1)
        close(ft, ref["onestep_up_ft1"], "OneStepSDPipeline up_ft[1]")
        close(ft.mean(0, keepdim=True), ref["onestep_mean"], "SDFeaturizer ensemble mean")


def test_typicality_loop_matches_reference_execution(ref, gen, unet_weights, vae_weights):
    """D.compute_losses (compute.py:134-160) executed by the reference vs the restated loop used by every parity test:
    grid[s, c] = (unet(add_noise(x, eps_s, t_s), t_s, ctx_c) - eps_s)^2, fp16, with the draws of D.noising"""
    from diff_mining_b200 import typicality as ty   # host-side mirror: load_image / draws replicate the reference RNG calls

    _, _, ctx, _ = gen.unet_inputs()
    embeds = torch.stack([ctx[0], ctx[1], ctx[0] * 0.5])
    pil = gen.pil_image()
    sd_stub = type("S", (), {"device": torch.device("cpu"), "scheduler": type("C", (), {"num_train_timesteps": 1000})()})()
    d = ty.D(sd_stub, "/tmp/unused", "cars", seed=42, N=5, t_min=0.1, t_max=0.7)
    x_img = d.load_image(pil)
    with torch.no_grad():
        torch.manual_seed(7)
        mean, logvar = sd15.vae_encode_moments(vae_weights, x_img)
        x = sd15.vae_sample(mean, logvar, torch.randn(mean.shape))
        close(x, ref["latent"], "encode_vae latent")
        noises, ts = d.draws(x)
        assert np.array_equal(ts.numpy(), ref["draw_t"]) and np.array_equal(noises.numpy(), ref["draw_noise"])
        grid = torch.empty(5, 3, 4, 6, 8)
        for c in range(3):
            pred = sd15.unet_forward(unet_weights, sd15.add_noise(x.expand(5, -1, -1, -1), noises, ts), ts, embeds[c][None].expand(5, -1, -1))
            grid[:, c] = (pred - noises) ** 2
        g16, r16 = grid.half().float(), torch.from_numpy(ref["losses_grid"]).float()
        assert ref["losses_grid"].dtype == np.float16
        # fp16 payload: allow one fp16 ulp where fp32 rounding noise crosses a rounding boundary
        assert ((g16 - r16).abs() <= 1e-3 * r16.abs() + 1e-6).all()
        assert (g16 == r16).float().mean() > 0.95
        close(grid[:2, 1], ref["compute_loss_rows"], "SD.compute_loss rows")


def test_consumers_match_reference_execution(ref, gen):
    cg = gen.consumer_grid()
    H, W, kx, ky = 48, 64, 16, 16
    Dm = consumers.patch_scores(cg, H, W, kx, ky)
    close(Dm, ref["D_map"], "load_typicality D map", rtol=1e-5, atol=1e-6)
    top = consumers.non_overlapping_topk(np.asarray(ref["D_map"]), kx, ky, 5)
    assert [list(r[:4]) for r in top] == ref["topk_boxes"].tolist()
    np.testing.assert_allclose([r[4] for r in top], ref["topk_scores"], rtol=1e-6)
    # T(x|c) as load_typicality_norm reduces it (before `normalize`): our host-side typicality_map
    from diff_mining_b200 import typicality as ty

    T = ty.typicality_map(torch.from_numpy(cg), size=(H, W)).numpy().astype(np.float64)
    # the display normalisation load_typicality_norm applies last (cluster.py:42-46): negatives / |min|, positives / max
    Tn = T.copy()
    Tn[T < 0] = T[T < 0] / np.abs(T.min())
    Tn[T > 0] = T[T > 0] / T.max()
    np.testing.assert_allclose((Tn + 1) / 2.0, np.asarray(ref["T_norm"], dtype=np.float64), rtol=1e-4, atol=1e-5)
