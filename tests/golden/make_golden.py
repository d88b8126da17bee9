"""Generates tests/golden/oracle_golden.npz.  The reference itself cannot be imported here (diffusers/xformers are
not installable offline), so these vectors are outputs of the ORACLE on seeded synthetic weights and
inputs, run in this container on CPU in fp32.  They pin the oracle against drift and give the GPU tests a
fixture that does not depend on the oracle code being importable on the GPU box.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sd15  # noqa: E402


def inputs():
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(2, 4, 16, 16, generator=g)
    t = torch.tensor([137, 642])
    ctx = torch.randn(2, 77, 768, generator=g)
    img = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    x_odd = torch.randn(1, 4, 9, 13, generator=g)
    return x, t, ctx, img, x_odd


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    vsd = sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1)
    x, t, ctx, img, x_odd = inputs()
    with torch.no_grad():
        eps = sd15.unet_forward(usd, x, t, ctx)
        feat = sd15.unet_forward(usd, x, t, ctx, up_ft_index=1)
        eps_odd = sd15.unet_forward(usd, x_odd, t[:1], ctx[:1])
        mean, logvar = sd15.vae_encode_moments(vsd, img)
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"),
        eps=eps.numpy(), feat_mean=feat.mean(dim=(2, 3)).numpy(), feat_corner=feat[:, :8, :4, :4].numpy(),
        eps_odd=eps_odd.numpy(), vae_mean=mean.numpy(), vae_logvar=logvar.numpy(),
        w_probe=np.array([usd["conv_in.weight"].flatten()[:8].numpy(), usd["mid_block.resnets.0.conv1.weight"].flatten()[:8].numpy()]),
    )
    print("written", eps.shape, feat.shape, eps_odd.shape, mean.shape)


if __name__ == "__main__":
    main()
