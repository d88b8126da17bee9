"""On-disk formats, CPU part: the diffusers pipeline directory the reference loads
(compute.py:65-70, 383-385; written by finetuning/base.py:245-259) -- safetensors fp32 / `.fp16.` variants, deprecated VAE
attention key names, packed-cache keys -- and the asynchronous `.npy` writer's byte-identity with np.save."""
import json
import os
import tempfile

import numpy as np
import pytest
import torch

from diff_mining_b200 import typicality as ty
from oracle import sd15


def write_diffusers_dir(root, usd, vsd, *, fp16=False, deprecated_vae=False, conv_attn=False):
    """synthetic pipeline directory in the diffusers layout: model_index.json, unet/, vae/ (+ configs)"""
    from safetensors.torch import save_file

    os.makedirs(os.path.join(root, "unet"))
    os.makedirs(os.path.join(root, "vae"))
    json.dump({"_class_name": "StableDiffusionPipeline", "unet": ["diffusers", "UNet2DConditionModel"],
               "vae": ["diffusers", "AutoencoderKL"]}, open(os.path.join(root, "model_index.json"), "w"))
    name = "diffusion_pytorch_model.fp16.safetensors" if fp16 else "diffusion_pytorch_model.safetensors"
    cast = (lambda t: t.half().contiguous()) if fp16 else (lambda t: t.float().contiguous())
    save_file({k: cast(v) for k, v in usd.items()}, os.path.join(root, "unet", name))
    v = {}
    ren = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
    for k, t in vsd.items():
        if deprecated_vae and ".attentions." in k and "group_norm" not in k:
            for new, old in ren.items():
                if f".{new}." in k:
                    k = k.replace(f".{new}.", f".{old}.")
                    if conv_attn and k.endswith(".weight"):
                        t = t[:, :, None, None]
                    break
        v[k] = cast(t)
    v["decoder.conv_in.weight"] = cast(torch.zeros(4, 4, 3, 3))   # the loader must ignore the decoder / post_quant_conv
    v["post_quant_conv.weight"] = cast(torch.zeros(4, 4, 1, 1))
    save_file(v, os.path.join(root, "vae", name))


@pytest.mark.parametrize("fp16,deprecated,conv", [(False, False, False), (True, True, False), (False, True, True)])
def test_load_diffusers_dir_key_schema(unet_weights, vae_weights, fp16, deprecated, conv):
    with tempfile.TemporaryDirectory() as td:
        write_diffusers_dir(td, unet_weights, vae_weights, fp16=fp16, deprecated_vae=deprecated, conv_attn=conv)
        sds = ty.load_diffusers_dir(td)
        assert list(sds) == ["unet", "vae"]
        assert set(sds["unet"]) == set(sd15.unet_param_shapes()) and set(sds["vae"]) == set(sd15.vae_encoder_param_shapes())
        for k, shp in sd15.vae_encoder_param_shapes().items():
            assert tuple(sds["vae"][k].shape) == tuple(shp), k
        # synthetic weights are fp16-representable: both variants carry the same values
        for k in ("encoder.mid_block.attentions.0.to_q.weight", "encoder.mid_block.attentions.0.to_out.0.bias", "quant_conv.weight"):
            assert torch.equal(sds["vae"][k].float(), vae_weights[k])
        assert torch.equal(sds["unet"]["conv_in.weight"].float(), unet_weights["conv_in.weight"])
        assert len(ty.weight_files(td)) == 2


def test_missing_directory_and_hub_id_fail_loudly():
    with pytest.raises(FileNotFoundError):
        ty.load_diffusers_dir("/nonexistent/dir")
    with pytest.raises(FileNotFoundError, match="local Hugging Face cache"):
        ty.load_diffusers_dir("runwayml/stable-diffusion-v1-5")   # no network, not cached: error, never a download
    with tempfile.TemporaryDirectory() as td:
        with pytest.raises(FileNotFoundError):
            ty.load_diffusers_dir(td)


def test_packed_cache_key_follows_the_checkpoint():
    from safetensors.torch import save_file

    with tempfile.TemporaryDirectory() as td:
        for sub in ("unet", "vae"):
            os.makedirs(os.path.join(td, "m", sub))
            save_file({"a": torch.zeros(8)}, os.path.join(td, "m", sub, "diffusion_pytorch_model.safetensors"))
        open(os.path.join(td, "m", "model_index.json"), "w").write("{}")
        k1 = ty.packed_cache_path(os.path.join(td, "m"), cache_dir=td)
        assert k1 == ty.packed_cache_path(os.path.join(td, "m"), cache_dir=td) and k1.endswith(".dmpk")
        save_file({"a": torch.ones(8)}, os.path.join(td, "m", "unet", "diffusion_pytorch_model.safetensors"))
        assert ty.packed_cache_path(os.path.join(td, "m"), cache_dir=td) != k1


def test_async_npy_writer_is_byte_identical_to_np_save():
    w = ty.AsyncNpyWriter("cpu", depth=2)
    with tempfile.TemporaryDirectory() as td:
        g = torch.Generator().manual_seed(0)
        arrs = [torch.randn(5 + (i % 2), 2, 4, 6, 8, generator=g).half() for i in range(9)]
        for i, a in enumerate(arrs):
            w.submit(os.path.join(td, "out", f"{i}.npy"), a)
        w.flush()
        for i, a in enumerate(arrs):
            ref = os.path.join(td, f"ref{i}.npy")
            np.save(open(ref, "wb"), a.numpy())      # the reference's call (compute.py:192)
            assert open(ref, "rb").read() == open(os.path.join(td, "out", f"{i}.npy"), "rb").read()
        w.submit(os.path.join(td, "nonexistent", "\0bad", "x.npy"), arrs[0])
        with pytest.raises(BaseException):
            w.flush()
    w.close()
