"""Python handle on the CUDA engine (include/dm_abi.h).  torch is used only for device memory, streams and
host<->device copies; every FLOP of the hot path runs in libdm_b200.so."""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _abi

MAX_CTX_SLOTS = 64
_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


class ContextSlots:
    """The engine's ONE allocator of text-context slots (its cross-attention K/V caches hold MAX_CTX_SLOTS contexts).

    Contexts are keyed by CONTENT (hash of the fp32 [77,768] values), so the same prompt embedding always maps to the same
    slot no matter which object (typicality.SD, dift.SDFeaturizer) or tensor storage it arrives through.  `acquire` takes
    every context one engine call needs and returns their slots; slots holding one of those contexts are pinned for the call,
    the least recently used other slot is recycled for each context that is not resident, and a call that needs more distinct
    contexts than there are slots raises.  Uploads (16 small K/V GEMMs per context) are stream-ordered with the call that
    follows.  Slots written through Engine.set_context() directly are "manual": the allocator never recycles them."""

    def __init__(self, engine: "Engine", n_slots: int = MAX_CTX_SLOTS):
        self.engine = engine
        self.n = n_slots
        self.key_of: List[Optional[bytes]] = [None] * n_slots
        self.manual = [False] * n_slots
        self.last_use = [0] * n_slots
        self.slot_of: Dict[bytes, int] = {}
        self.tick = 0
        self.uploads = 0

    @staticmethod
    def key(ctx: torch.Tensor) -> bytes:
        import hashlib

        a = ctx.detach().to(device="cpu", dtype=torch.float32).contiguous()
        if tuple(a.shape) != (77, 768):
            raise ValueError(f"context must be [77, 768], got {tuple(a.shape)}")
        return hashlib.blake2b(a.numpy().tobytes(), digest_size=16).digest()

    def mark_manual(self, slot: int) -> None:
        k = self.key_of[slot]
        if k is not None:
            self.slot_of.pop(k, None)
        self.key_of[slot] = None
        self.manual[slot] = True

    def acquire(self, ctxs: Sequence[torch.Tensor]) -> List[int]:
        keys = [self.key(c) for c in ctxs]
        need = {}
        for k, c in zip(keys, ctxs):
            need.setdefault(k, c)
        free = [s for s in range(self.n) if not self.manual[s]]
        if len(need) > len(free):
            raise RuntimeError(f"one call needs {len(need)} distinct text contexts but the engine has {len(free)} context slots "
                               f"({self.n} total, {self.n - len(free)} set manually)")
        self.tick += 1
        for k in need:
            if k in self.slot_of:
                self.last_use[self.slot_of[k]] = self.tick
        for k, c in need.items():
            if k in self.slot_of:
                continue
            # victim: an unused slot first, else the least recently used slot that this call does not need
            cand = [s for s in free if self.key_of[s] is None or self.key_of[s] not in need]
            s = min(cand, key=lambda i: (self.key_of[i] is not None, self.last_use[i], i))
            old = self.key_of[s]
            if old is not None:
                del self.slot_of[old]
            self.engine._upload_context(s, c)
            self.uploads += 1
            self.key_of[s], self.slot_of[k], self.last_use[s] = k, s, self.tick
        return [self.slot_of[k] for k in keys]


class Engine:
    """One engine per (process, GPU).  Not re-entrant; all work is enqueued on torch's current stream."""

    def __init__(self, device: int | torch.device | str = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("dm_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.lib = _abi.load()
        h = ctypes.c_void_p()
        _abi.check(self.lib.dm_create(self.device.index or 0, ctypes.byref(h)))
        self._h = h
        self._finalized = False
        self.contexts = ContextSlots(self)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dm_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def patch_topk(self, T: torch.Tensor, H: int, W: int, kx: int = 64, ky: int = 64, k: int = 5, descending: bool = True):
        """The reference's T-map consumer on the GPU (cluster.py:125-137,183-205; utils.py:74-102): T [B,h,w] fp32
        (one condition of `typicality()`'s T output) -> bilinear resize to H x W -> kx x ky average pooling -> the k
        best mutually non-overlapping windows.  Returns (boxes int32 [B,k,4] = x_start, y_start, x_end, y_end in the
        reference's (row, column) convention, scores fp32 [B,k], count int32 [B]) on the device."""
        T = T.to(self.device, torch.float32).contiguous()
        B, h, w = T.shape
        work = torch.empty(int(self.lib.dm_patch_topk_work_floats(B, H, W, ky)), device=self.device, dtype=torch.float32)
        boxes = torch.zeros(B, k, 4, device=self.device, dtype=torch.int32)
        scores = torch.zeros(B, k, device=self.device, dtype=torch.float32)
        count = torch.zeros(B, device=self.device, dtype=torch.int32)
        _abi.check(self.lib.dm_patch_topk(_ptr(T), B, h, w, H, W, kx, ky, k, 1 if descending else 0, _ptr(work), _ptr(boxes),
                                          _ptr(scores), _ptr(count), self._stream()))
        return boxes, scores, count

    def set_variant(self, name: str, value: int) -> None:
        """kernel-variant switch for tests / A-B timing (process-wide; -1 = default): "igemm_pair", "gn_fused", "xattn",
        "prefix_share" (include/dm_abi.h: dm_op_set_variant).  Plans built earlier keep the variant they were built with."""
        _abi.check(self.lib.dm_op_set_variant(name.encode(), int(value)))

    # ---------------------------------------------------------------- weights / context / schedule
    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str) -> None:
        """prefix is 'unet.' or 'vae.' ; keys follow the diffusers schema."""
        for k, v in sd.items():
            t = v.detach().to("cpu").contiguous()
            if t.dtype not in _DT:
                t = t.float()
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            _abi.check(self.lib.dm_load_tensor(self._h, (prefix + k).encode(), ctypes.c_void_p(t.data_ptr()), _DT[t.dtype],
                                               t.dim(), shape))

    def finalize(self) -> None:
        _abi.check(self.lib.dm_finalize_weights(self._h))
        self._finalized = True

    def save_packed(self, path: str) -> None:
        """packed-weight cache: the engine's packed device buffers -> one file"""
        _abi.check(self.lib.dm_save_packed(self._h, os.fsencode(path)))

    def load_packed(self, path: str) -> None:
        """fresh engine <- packed-weight cache (replaces load_state_dict + finalize)"""
        _abi.check(self.lib.dm_load_packed(self._h, os.fsencode(path)))
        self._finalized = True

    def set_schedule(self, sqrt_acp: torch.Tensor, sqrt_1m_acp: torch.Tensor) -> None:
        a = sqrt_acp.detach().float().cpu().contiguous()
        b = sqrt_1m_acp.detach().float().cpu().contiguous()
        _abi.check(self.lib.dm_set_schedule(self._h, _ptr(a), _ptr(b), a.numel()))

    def set_context(self, slot: int, ctx: torch.Tensor) -> None:
        """write a context into an explicit slot (low-level use: tests, bench); the slot leaves the allocator's pool"""
        self._upload_context(slot, ctx)
        self.contexts.mark_manual(int(slot))

    def _upload_context(self, slot: int, ctx: torch.Tensor) -> None:
        c = ctx.detach().float().cpu().contiguous()
        if tuple(c.shape) != (77, 768):
            raise ValueError(f"context must be [77, 768], got {tuple(c.shape)}")
        _abi.check(self.lib.dm_set_context(self._h, int(slot), _ptr(c), self._stream()))

    # ---------------------------------------------------------------- hot path
    def vae_encode(self, img: torch.Tensor, eps: Optional[torch.Tensor] = None, return_moments: bool = False):
        img = _dev_f32(img, self.device)
        B, C, H, W = img.shape
        if C != 3:
            raise ValueError("vae_encode expects [B,3,H,W]")
        h, w = H // 8, W // 8
        z = torch.empty(B, 4, h, w, device=self.device, dtype=torch.float32)
        mean = torch.empty_like(z) if return_moments else None
        logvar = torch.empty_like(z) if return_moments else None
        e = None if eps is None else _dev_f32(eps, self.device)
        _abi.check(self.lib.dm_vae_encode(self._h, _ptr(img), _ptr(e), B, H, W, _ptr(z), _ptr(mean), _ptr(logvar),
                                          self._stream()))
        return (z, mean, logvar) if return_moments else z

    def unet_eps(self, x_noisy: torch.Tensor, t: torch.Tensor, ctx_slots: Sequence[int]) -> torch.Tensor:
        x = _dev_f32(x_noisy, self.device)
        Bf, _, h, w = x.shape
        tt = t.to(device=self.device, dtype=torch.int64).expand(Bf).contiguous()
        out = torch.empty_like(x)
        s = _i32(ctx_slots)
        assert s.shape[0] == Bf
        _abi.check(self.lib.dm_unet_eps(self._h, _ptr(x), _ptr(tt), s.ctypes.data_as(ctypes.c_void_p), Bf, h, w, _ptr(out),
                                        self._stream()))
        return out

    def unet_rows(self, x: torch.Tensor, noise: Optional[torch.Tensor], t: torch.Tensor, x_index, noise_index, ctx_slots,
                  want_loss: bool = True, want_eps: bool = False, max_forwards: int = 0):
        x = _dev_f32(x, self.device)
        n = None if noise is None else _dev_f32(noise, self.device)
        tt = t.to(device=self.device, dtype=torch.int64).contiguous()
        cs = _i32(ctx_slots)
        M = cs.shape[0]
        h, w = x.shape[-2:]
        xi = None if x_index is None else _i32(x_index)
        ni = None if noise_index is None else _i32(noise_index)
        loss = torch.empty(M, 4, h, w, device=self.device, dtype=torch.float32) if want_loss else None
        eps = torch.empty(M, 4, h, w, device=self.device, dtype=torch.float32) if want_eps else None
        vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _abi.check(self.lib.dm_unet_rows(self._h, _ptr(x), _ptr(n), _ptr(tt), vp(xi), vp(ni), vp(cs), M, h, w, _ptr(loss),
                                         _ptr(eps), max_forwards, self._stream()))
        return loss, eps

    def typicality(self, x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, ctx_slots: Sequence[int],
                   want_grid: bool = True, want_T: bool = True, max_forwards: int = 0):
        """x0 [Bi,4,h,w], noise [N,4,h,w], t [N]; ctx_slots = condition slots with the unconditional LAST.
        Returns (grid fp16 [Bi,N,n_cond,4,h,w] | None, T fp32 [Bi,n_cond-1,h,w] | None)."""
        x0 = _dev_f32(x0, self.device)
        noise = _dev_f32(noise, self.device)
        tt = t.to(device=self.device, dtype=torch.int64).contiguous()
        Bi, _, h, w = x0.shape
        N = noise.shape[0]
        cs = _i32(ctx_slots)
        n_cond = cs.shape[0]
        grid = torch.empty(Bi, N, n_cond, 4, h, w, device=self.device, dtype=torch.float16) if want_grid else None
        T = torch.empty(Bi, n_cond - 1, h, w, device=self.device, dtype=torch.float32) if want_T else None
        _abi.check(self.lib.dm_typicality(self._h, _ptr(x0), _ptr(noise), _ptr(tt), cs.ctypes.data_as(ctypes.c_void_p), Bi,
                                          N, n_cond, h, w, _ptr(grid), _ptr(T), max_forwards, self._stream()))
        return grid, T

    def dift(self, latents: torch.Tensor, noise: Optional[torch.Tensor], t: int, ctx_slot: int, ensemble: int,
             up_ft_index: int = 1) -> torch.Tensor:
        lat = _dev_f32(latents, self.device)
        BE, _, h, w = lat.shape
        if BE % ensemble:
            raise ValueError("latents batch must be a multiple of the ensemble size")
        B = BE // ensemble
        C, ho, wo = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _abi.check(self.lib.dm_dift_shape(h, w, up_ft_index, ctypes.byref(C), ctypes.byref(ho), ctypes.byref(wo)))
        out = torch.empty(B, C.value, ho.value, wo.value, device=self.device, dtype=torch.float32)
        n = None if noise is None else _dev_f32(noise, self.device)
        _abi.check(self.lib.dm_dift(self._h, _ptr(lat), _ptr(n), int(t), int(ctx_slot), B, ensemble, h, w, up_ft_index,
                                    _ptr(out), self._stream()))
        return out

    # ---------------------------------------------------------------- introspection
    @property
    def launch_count(self) -> int:
        return int(self.lib.dm_launch_count(self._h))

    @property
    def flop_count(self) -> float:
        return float(self.lib.dm_flop_count(self._h))

    def debug_keep(self, on: bool) -> None:
        _abi.check(self.lib.dm_debug_keep(self._h, int(on)))

    def debug_fetch(self, name: str) -> torch.Tensor:
        dims = (ctypes.c_int * 4)()
        n = self.lib.dm_debug_fetch(self._h, name.encode(), None, 0, dims, self._stream())
        if n < 0:
            raise RuntimeError("dm_b200: " + self.lib.dm_last_error().decode())
        out = torch.empty(tuple(dims), device=self.device, dtype=torch.float32)
        n = self.lib.dm_debug_fetch(self._h, name.encode(), _ptr(out), out.numel(), dims, self._stream())
        if n < 0:
            raise RuntimeError("dm_b200: " + self.lib.dm_last_error().decode())
        torch.cuda.synchronize(self.device)
        return out

    def profile_plan(self, kind: str, Bf: int, h: int, w: int, aux: int = 0, iters: int = 3) -> dict:
        """CUDA-event timing of one replay of a cached plan by op class (roofline reporting).  kind: "unet" (aux = rows per
        shared-prefix group), "dift" (aux = up_ft_index) or "vae" (h, w = image size)."""
        ms = (ctypes.c_double * 3)()
        fl = (ctypes.c_double * 2)()
        _abi.check(self.lib.dm_profile_plan(self._h, {"unet": 0, "dift": 1, "vae": 2}[kind], Bf, h, w, aux, iters, ms, fl))
        return {"ms_igemm": ms[0], "ms_attn": ms[1], "ms_other": ms[2], "flops_igemm": fl[0], "flops_attn": fl[1]}

    def profile_unet(self, Bf: int, h: int, w: int, iters: int = 3) -> dict:
        v = [ctypes.c_double() for _ in range(5)]
        _abi.check(self.lib.dm_profile_unet(self._h, Bf, h, w, iters, *[ctypes.byref(x) for x in v]))
        return {"ms_igemm": v[0].value, "ms_attn": v[1].value, "ms_other": v[2].value, "flops_igemm": v[3].value,
                "flops_attn": v[4].value}
