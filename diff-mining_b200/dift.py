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
from .typicality import VAE_SCALING, load_diffusers_dir, scaled_linear_schedule


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
            sds = state_dicts if state_dicts is not None else load_diffusers_dir(sd_id)
            engine.load_state_dict(sds["unet"], "unet.")
            engine.load_state_dict(sds["vae"], "vae.")
            engine.finalize()
            engine.set_schedule(*scaled_linear_schedule())
        self.engine = engine
        self._embeds = dict(prompt_embeds or {})
        self._text_encoder = None
        self._text_encoder_id = text_encoder_id or "openai/clip-vit-large-patch14"
        self._slots: Dict[str, int] = {}
        self._next_slot = 48  # keep clear of the category slots typicality.SD assigns from 0

    def _slot_for(self, prompt: Union[str, torch.Tensor]) -> int:
        key = prompt if isinstance(prompt, str) else f"tensor@{prompt.data_ptr()}"
        if key not in self._slots:
            if isinstance(prompt, torch.Tensor):
                emb = prompt.reshape(77, 768)
            elif prompt in self._embeds:
                emb = self._embeds[prompt]
            else:
                if self._text_encoder is None:
                    from .text import ClipTextEncoder

                    self._text_encoder = ClipTextEncoder(self._text_encoder_id, self.device)
                emb = self._text_encoder([prompt])[0]
            s = self._next_slot
            self._next_slot = 48 + (self._next_slot - 48 + 1) % 16
            for k in [k for k, v in self._slots.items() if v == s]:
                del self._slots[k]
            self.engine.set_context(s, emb)
            self._slots[key] = s
        return self._slots[key]

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
