"""Text context for the engine: CLIP text tower -> last_hidden_state [n,77,768] fp32, as
CategoryFeatures.embed does (/root/reference/diffmining/typicality/compute.py:28-51).  Per SURVEY.md R10 the text
encoder is a one-off start-up step that stays on transformers/torch -- the engine only ever sees [77,768] tensors."""
from __future__ import annotations

from typing import List

import torch


class ClipTextEncoder:
    def __init__(self, name: str, device, dtype=torch.float16):
        from transformers import CLIPTextModel, CLIPTokenizer

        try:
            self.tokenizer = CLIPTokenizer.from_pretrained(name, local_files_only=True)
            self.model = CLIPTextModel.from_pretrained(name, torch_dtype=dtype, local_files_only=True).to(device).eval()
        except Exception as ex:  # offline box without the checkpoint: fail loudly, never substitute embeddings
            raise RuntimeError(
                f"CLIP text encoder '{name}' is not available locally ({type(ex).__name__}: {ex}). "
                "Pass precomputed `category_embeds` ({category: [77,768]}) to SD(...) instead.") from ex
        self.device = device

    @torch.no_grad()
    def __call__(self, prompts: List[str]) -> torch.Tensor:
        ids = self.tokenizer(prompts, max_length=self.tokenizer.model_max_length, padding="max_length", truncation=True,
                             return_tensors="pt").input_ids
        return self.model(ids.to(self.device))[0].float()
