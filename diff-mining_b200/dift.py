"""Drop-in for the reference's DIFT featurizer -- `SDFeaturizer.forward` of
/root/reference/diffmining/typicality/dift.py:195-232 -- on top of the CUDA engine.

forward(img_tensor[3,H,W] | [1,3,H,W], prompt, t=261, up_ft_index=1, ensemble_size=8) -> [1, C, H/16.., W/16..]
(on device), the mean over `ensemble_size` noise members of the U-Net activation after `up_blocks[up_ft_index]`.

The reference repeats the image `ensemble_size` times and VAE-encodes every copy (dift.py:187,220); the copies differ
only by the posterior draw, so the encoder runs ONCE here and the `ensemble_size` posterior samples are formed from the
same moments -- identical results for identical draws."""
from __future__ import annotations

from typing import Dict, Optional, Union

import numpy as np
import torch

from .engine import Engine
from .typicality import VAE_SCALING, load_engine_weights, scaled_linear_schedule


def dift_pre(img) -> torch.Tensor:
    """PIL -> [3,H,W] in [-1,1] (dift.py:19-21)"""
    a = torch.from_numpy(np.asarray(img.convert("RGB"), dtype=np.uint8).copy()).permute(2, 0, 1).float()
    return (a / 255.0 - 0.5) * 2


class SDFeaturizer(object):
    def __init__(self, sd_id, base_id="runwayml/stable-diffusion-v1-5", text_encoder_id=None, *, engine: Optional[Engine] = None,
                 state_dicts=None, prompt_embeds: Optional[Dict[str, torch.Tensor]] = None, device="cuda:0"):
        """`engine` shares an already-loaded Engine (e.g. SD(...).engine); `state_dicts` / `prompt_embeds` allow
        offline construction exactly like typicality.SD."""
        self.device = torch.device(device)
        if engine is None:
            engine = Engine(self.device)
            load_engine_weights(engine, sd_id, state_dicts)
            engine.set_schedule(*scaled_linear_schedule())
        self.engine = engine
        self._embeds = dict(prompt_embeds or {})
        self._text_encoder = None
        self._text_encoder_id = text_encoder_id or "openai/clip-vit-large-patch14"

    def _embed(self, prompt: Union[str, torch.Tensor]) -> torch.Tensor:
        if isinstance(prompt, torch.Tensor):
            return prompt.reshape(77, 768)
        if prompt not in self._embeds:
            if self._text_encoder is None:
                from .text import ClipTextEncoder

                self._text_encoder = ClipTextEncoder(self._text_encoder_id, self.device)
            self._embeds[prompt] = self._text_encoder([prompt])[0]
        return self._embeds[prompt]

    def _slot_for(self, prompt: Union[str, torch.Tensor]) -> int:
        """context slot of a prompt through the engine's shared, content-keyed allocator (engine.ContextSlots): safe next to
        typicality.SD on the same engine, and tensor prompts are identified by their values, not their storage"""
        return self.engine.contexts.acquire([self._embed(prompt)])[0]

    @torch.no_grad()
    def forward(self, img_tensor, prompt, t=261, up_ft_index=1, ensemble_size=8):
        if img_tensor.dim() == 3:
            img_tensor = img_tensor.unsqueeze(0)
        img_tensor = img_tensor.to(self.device).float()
        B = img_tensor.shape[0]
        slot = self._slot_for(prompt)
        _, mean, logvar = self.engine.vae_encode(img_tensor, None, return_moments=True)
        # posterior samples + forward noise, drawn with torch as the reference does (dift.py:187-189)
        shape = (B, ensemble_size) + tuple(mean.shape[1:])
        post = torch.randn(shape, device=self.device, dtype=torch.float32)
        std = torch.exp(0.5 * logvar)
        latents = ((mean[:, None] + std[:, None] * post) * VAE_SCALING).reshape((B * ensemble_size,) + tuple(mean.shape[1:]))
        noise = torch.randn_like(latents)
        return self.engine.dift(latents, noise, int(t), slot, ensemble_size, up_ft_index)

    __call__ = forward

    # ------------------------------------------------------------------ caller-side dedupe
    @torch.no_grad()
    def forward_many(self, images, prompt, t=261, up_ft_index=1, ensemble_size=8, cache=True):
        """Feature maps of MANY images: list of [3,H,W] / [1,3,H,W] tensors -> list of [1, C, h, w] maps, each identical to
        `forward(image, ...)` given the same torch RNG state at entry (the posterior and forward-noise draws are made image
        by image in list order, exactly the draws the per-image calls would make).

        What the reference's caller does per PATCH (cluster.py:255-299: `self.embed(image, prompt, t)` inside the loop over
        the <= 5 windows of an image, each call encoding the image `ensemble_size` times, dift.py:187,220) is done here once
        per IMAGE: one VAE encode, one `ensemble_size`-member partial forward, maps cached by image content, and images of
        equal size share engine launches of 64 members (dm_dift's micro-batch)."""
        if not hasattr(self, "_fmap_cache"):
            self._fmap_cache = {}
        imgs = [(im.unsqueeze(0) if im.dim() == 3 else im).to(self.device).float() for im in images]
        pkey = ContextKey.of(self._embed(prompt)) + bytes(f"|{int(t)}|{int(up_ft_index)}|{int(ensemble_size)}", "ascii")
        keys = [pkey + ContextKey.of_image(im) for im in imgs] if cache else [None] * len(imgs)
        out = [self._fmap_cache.get(k) if cache else None for k in keys]
        todo = [i for i, o in enumerate(out) if o is None]
        # an image that appears twice in one call is computed once
        first_of = {}
        for i in list(todo):
            if cache and keys[i] in first_of:
                todo.remove(i)
            elif cache:
                first_of[keys[i]] = i
        slot = self._slot_for(prompt)
        E = ensemble_size
        # draws in list order (RNG parity with per-image forward() calls), then batches by image size
        prepared = {}
        for i in todo:
            _, mean, logvar = self.engine.vae_encode(imgs[i], None, return_moments=True)
            post = torch.randn((1, E) + tuple(mean.shape[1:]), device=self.device, dtype=torch.float32)
            lat = ((mean[:, None] + torch.exp(0.5 * logvar)[:, None] * post) * VAE_SCALING).reshape((E,) + tuple(mean.shape[1:]))
            prepared[i] = (lat, torch.randn_like(lat))
        by_shape = {}
        for i in todo:
            by_shape.setdefault(tuple(prepared[i][0].shape[1:]), []).append(i)
        for shape, idx in by_shape.items():
            lat = torch.cat([prepared[i][0] for i in idx])
            nz = torch.cat([prepared[i][1] for i in idx])
            ft = self.engine.dift(lat, nz, int(t), slot, E, up_ft_index)   # [len(idx), C, h, w]; 64 members per launch
            for k, i in enumerate(idx):
                out[i] = ft[k:k + 1]
                if cache:
                    self._fmap_cache[keys[i]] = out[i]
        if cache:
            for i, k in enumerate(keys):
                if out[i] is None:
                    out[i] = self._fmap_cache[k]
        return out

    def clear_cache(self):
        self._fmap_cache = {}

    @torch.no_grad()
    def patch_descriptors(self, image, boxes, prompt, t=261, up_ft_index=1, ensemble_size=8):
        """L2-normalised DIFT descriptors of the windows `boxes` = [(x_start, y_start, x_end, y_end), ...] (the reference's
        (row, column) convention) of one image, cropped from the image's cached feature map exactly as
        cluster.py:283-299 does: emb[:, int(x_start*H):int(x_end*H), int(y_start*W):int(y_end*W)].mean((1,2)) / norm."""
        im = image.unsqueeze(0) if image.dim() == 3 else image
        fmap = self.forward_many([im], prompt, t, up_ft_index, ensemble_size)[0][0]   # [C, h, w]
        _, h, w = fmap.shape
        Hs, Ws = h / im.shape[-2], w / im.shape[-1]
        out = []
        for (x0, y0, x1, y1) in boxes:
            e = fmap[:, int(x0 * Hs):int(x1 * Hs), int(y0 * Ws):int(y1 * Ws)].mean(dim=(1, 2))
            out.append(e / torch.linalg.vector_norm(e))
        return torch.stack(out)


class ContextKey:
    """content hashes used as cache keys"""

    @staticmethod
    def of(t: torch.Tensor) -> bytes:
        import hashlib

        return hashlib.blake2b(t.detach().float().cpu().contiguous().numpy().tobytes(), digest_size=16).digest()

    @staticmethod
    def of_image(t: torch.Tensor) -> bytes:
        import hashlib

        a = t.detach().contiguous()
        h = hashlib.blake2b(digest_size=16)
        h.update(str(tuple(a.shape)).encode())
        h.update(a.cpu().numpy().tobytes())
        return h.digest()
