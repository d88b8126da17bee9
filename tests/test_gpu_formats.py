"""GPU tests of the rows SURVEY.md 8f-2 / 8f-3 and of the ADVICE r01 items: diffusers-directory load through SD(model_path=...),
packed-weight cache, asynchronous .npy writer, SDFeaturizer.forward_many, the shared context-slot allocator."""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _embeds(contexts, n_extra=0):
    e = {"": contexts[0], "1975": contexts[1], "1995": contexts[2]}
    g = torch.Generator().manual_seed(17)
    for i in range(n_extra):
        e[f"c{i:03d}"] = torch.randn(77, 768, generator=g)
    return e


def test_sd_from_diffusers_directory_and_packed_cache(unet_weights, vae_weights, contexts, engine):
    """SD(model_path=<diffusers dir>) == SD(state_dicts=...) bit for bit, for the fp32 layout with the deprecated VAE
    attention names stored as 1x1 convs and for the `.fp16.` variant; the second construction comes from the packed cache"""
    from diff_mining_b200.typicality import SD
    from tests.test_formats import write_diffusers_dir

    g = torch.Generator().manual_seed(4)
    img = torch.rand(1, 3, 64, 96, generator=g) * 2 - 1
    x = torch.randn(2, 4, 8, 12, generator=g)
    t = torch.tensor([20, 800])
    ref_eps = engine.unet_eps(x, t, [1, 2])
    _, ref_mean, ref_lv = engine.vae_encode(img, None, return_moments=True)
    emb = _embeds(contexts)
    with tempfile.TemporaryDirectory() as td:
        os.environ["DM_WEIGHT_CACHE"] = os.path.join(td, "cache")
        try:
            for variant, kw in (("fp32", dict(deprecated_vae=True, conv_attn=True)), ("fp16", dict(fp16=True))):
                root = os.path.join(td, variant)
                write_diffusers_dir(root, unet_weights, vae_weights, **kw)
                for expect in ("diffusers", "packed_cache"):
                    sd = SD("cars", root, ["1975", "1995"], DEV, True, category_embeds=emb)
                    assert sd.weights_source == expect
                    slots = sd.slots_for(torch.stack([contexts[1], contexts[2]]))
                    assert torch.equal(sd.engine.unet_eps(x, t, slots), ref_eps)
                    _, m, lv = sd.engine.vae_encode(img, None, return_moments=True)
                    assert torch.equal(m, ref_mean) and torch.equal(lv, ref_lv)
                    sd.engine.close()
            assert len(os.listdir(os.path.join(td, "cache"))) == 2
        finally:
            del os.environ["DM_WEIGHT_CACHE"]


def test_async_writer_matches_synchronous_compute(unet_weights, vae_weights, contexts):
    """D.compute with async_write: byte-identical .npy files, written while the next image is being scored"""
    from PIL import Image

    from diff_mining_b200.typicality import D, SD

    sd = SD("cars", None, ["1975", "1995"], DEV, True, state_dicts={"unet": unet_weights, "vae": vae_weights},
            category_embeds=_embeds(contexts))
    rng = np.random.RandomState(1)
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for i in range(5):
            p = os.path.join(td, "src", f"1975__car_{i:03d}.png")
            os.makedirs(os.path.dirname(p), exist_ok=True)
            Image.fromarray(rng.randint(0, 255, (64 + 8 * (i % 2), 96, 3), dtype=np.uint8)).save(p)
            paths.append(p)
        d_sync = D(sd, os.path.join(td, "sync"), "geo", seed=42, N=4, t_min=0.1, t_max=0.7)
        d_async = D(sd, os.path.join(td, "async"), "geo", seed=42, N=4, t_min=0.1, t_max=0.7, async_write=True)
        for d in (d_sync, d_async):
            torch.manual_seed(3)   # the (unseeded) VAE posterior draws
            for p in paths:
                d.compute("1975", p)
            d.flush()
        for p in paths:
            a, b = open(d_sync.get_path(p), "rb").read(), open(d_async.get_path(p), "rb").read()
            assert a == b and len(a) > 1000
            assert d_async(p).dtype == np.float16
    sd.engine.close()


def test_forward_many_matches_per_image_forward(engine, contexts):
    """one VAE encode + one ensemble forward per image, batched across images, cached per image -- the same
    bits as per-image SDFeaturizer.forward calls with the same RNG stream; patch descriptors as cluster.py:283-299"""
    from diff_mining_b200.dift import SDFeaturizer

    f = SDFeaturizer(None, engine=engine, prompt_embeds={"a car": contexts[1]}, device=DEV)
    g = torch.Generator().manual_seed(9)
    imgs = [torch.rand(3, 64, 96, generator=g) * 2 - 1, torch.rand(3, 64, 96, generator=g) * 2 - 1, torch.rand(3, 128, 64, generator=g) * 2 - 1,
            torch.rand(3, 64, 96, generator=g) * 2 - 1]
    torch.manual_seed(5)
    single = [f.forward(im, "a car", t=161, up_ft_index=1, ensemble_size=4) for im in imgs]
    torch.manual_seed(5)
    launches0 = engine.launch_count
    many = f.forward_many(imgs, "a car", t=161, up_ft_index=1, ensemble_size=4, cache=False)
    batched_launches = engine.launch_count - launches0
    for a, b in zip(single, many):
        assert a.shape == b.shape and torch.equal(a, b)
    # cache: a second call computes nothing; a repeated image inside one call is computed once
    f.clear_cache()
    torch.manual_seed(5)
    first = f.forward_many([imgs[0], imgs[0], imgs[2]], "a car", t=161, up_ft_index=1, ensemble_size=4)
    assert torch.equal(first[0], first[1])
    l1 = engine.launch_count
    again = f.forward_many([imgs[2], imgs[0]], "a car", t=161, up_ft_index=1, ensemble_size=4)
    assert engine.launch_count == l1 and torch.equal(again[0], first[2]) and torch.equal(again[1], first[0])
    assert batched_launches > 0
    # descriptors of windows cropped from the cached map
    boxes = [(0, 0, 32, 32), (16, 40, 64, 96)]
    desc = f.patch_descriptors(imgs[0], boxes, "a car", t=161, up_ft_index=1, ensemble_size=4)
    fmap = first[0][0]
    H, W = fmap.shape[1] / 64, fmap.shape[2] / 96
    for k, (x0, y0, x1, y1) in enumerate(boxes):
        e = fmap[:, int(x0 * H):int(x1 * H), int(y0 * W):int(y1 * W)].mean(dim=(1, 2))
        torch.testing.assert_close(desc[k], e / e.norm(), rtol=1e-6, atol=1e-7)
    assert desc.shape == (2, 1280)


def test_more_categories_than_slots(unet_weights, contexts):
    """the reference passes every category to SD (365 for places); the drop-in must construct and score any
    (category, "") pair, and SD + SDFeaturizer on one engine must not overwrite each other's contexts"""
    from diff_mining_b200.dift import SDFeaturizer
    from diff_mining_b200.typicality import SD

    emb = _embeds(contexts, n_extra=100)
    cats = [c for c in emb if c]
    sd = SD("places", None, cats, DEV, True, state_dicts={"unet": unet_weights}, category_embeds=emb)
    assert len(sd.country_embeds) == 103
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 4, 8, 8, generator=g)
    noise = torch.randn(1, 4, 8, 8, generator=g)
    t = torch.tensor([300])
    ref = {}
    for c in ("c000", "c050", "c099", "1975"):
        ce = torch.stack([sd.country_embeds[c], sd.country_embeds[""]])
        ref[c] = sd.compute_loss(x, noise, t, ce).clone()
    feat = SDFeaturizer(None, engine=sd.engine, prompt_embeds={"p": contexts[2]}, device=DEV)
    for i in range(100):   # churn through far more contexts than slots
        ce = torch.stack([sd.country_embeds[f"c{i:03d}"], sd.country_embeds[""]])
        sd.slots_for(ce)
        feat._slot_for("p")
    for c, r in ref.items():
        ce = torch.stack([sd.country_embeds[c], sd.country_embeds[""]])
        assert torch.equal(sd.compute_loss(x, noise, t, ce), r)
    with pytest.raises(RuntimeError, match="distinct text contexts"):
        sd.slots_for(torch.stack([sd.country_embeds[c] for c in cats[:70]]))
    sd.engine.close()
