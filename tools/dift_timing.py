"""Timing of BASELINE config 3 (DIFT-161 features: 64 members = 8 images x ensemble 8 at 512x512, t = 161, up_ft_index 1),
with and without the VAE encode.  python tools/dift_timing.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diff_mining_b200.engine import Engine  # noqa: E402
from oracle import sd15  # noqa: E402  (synthetic weights only)

eng = Engine(0)
eng.load_state_dict(sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0), "unet.")
eng.load_state_dict(sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1), "vae.")
eng.finalize()
eng.set_schedule(*sd15.schedule_tables())
g = torch.Generator().manual_seed(5)
eng.set_context(0, torch.randn(77, 768, generator=g))
imgs = (torch.rand(8, 3, 512, 512, generator=g) * 2 - 1).cuda()
E = 8


def members():
    _, mean, logvar = eng.vae_encode(imgs, None, return_moments=True)
    post = torch.randn((8, E) + tuple(mean.shape[1:]), device="cuda")
    lat = ((mean[:, None] + torch.exp(0.5 * logvar)[:, None] * post) * 0.18215).reshape((8 * E,) + tuple(mean.shape[1:]))
    return lat, torch.randn_like(lat)


def timed(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


lat, noise = members()
ms_unet = timed(lambda: eng.dift(lat, noise, 161, 0, E, 1))
ms_all = timed(lambda: eng.dift(*members(), 161, 0, E, 1))
gf = 64 * 438.8
print(f"DIFT config 3: 64 members, partial U-Net only {ms_unet:.2f} ms ({64e3 / ms_unet:.0f} members/s, {gf / ms_unet:.0f} TFLOP/s); "
      f"with encode-once VAE (8 encodes) {ms_all:.2f} ms ({64e3 / ms_all:.0f} members/s)")
