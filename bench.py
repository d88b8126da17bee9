#!/usr/bin/env python
"""bench.py -- benchmarks of the typicality / DIFT hot path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|5]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              bench.py --gpus N --steps K --warmup W

--config 2 (default, the headline BASELINE.json metric, configs[1]): per GPU and step, 16 synthetic 512x512 RGB images,
    each scored with 32 (eps,t) draws x {c, uncond} = 64 SD-1.5 U-Net forwards at 64x64 latents + 1 VAE encode
    (52.5 TFLOP / image).  One "sample" = one image scored.
--config 3 (configs[2]): DIFT-161 features, t=161: per GPU and step 8 images 512x512 x ensemble 8 = 64 members
    (one VAE encode per image, 64 partial U-Net forwards through up_blocks[1]); unit = members/s.
--config 4 (configs[3], meant for --gpus 8): 10 000 synthetic 512x512 images in total, sharded over the ranks (rank r takes
    images r::world as compute.py:336-341 does), 10 "country" prompts + uncond resident as contexts, every image scored with 32
    draws x {its country, uncond}; per-image raw grids stay rank-local, the T maps of all 10 000 images are gathered with ONE
    all-gather at the end of the step (164 MB).  Fixed total work: "scaling": "strong".  One step = the whole data set.
--config 5 (configs[4]): 1024x1024 image, 14 conditions + uncond batched in one call, 16 (eps,t) draws = 240 U-Net
    forwards at 128x128 latents + 1 VAE encode per image (1126.6 TFLOP / image); one image per GPU and step.
Weak scaling everywhere: every rank works on its own inputs; for the typicality configs the only exchange is the
all-gather of the per-image T maps, issued asynchronously per step and completed inside the timed region.

Printed JSON (one line, rank 0):
  value   units/s, inputs resident in HBM (fp32 images on the device), outputs left on the device
  e2e     units/s through the host-facing path: pinned host images -> H2D -> engine -> D2H of the step's results
          (typicality: the raw fp16 loss grid -- the reference's .npy payload -- and the T maps; DIFT: the feature maps)
  roofline   the dominant kernel (tcgen05 implicit GEMM, all launches of one micro-batch of the benched plan):
          algorithmic FLOPs / CUDA-event time per launch, measured live (eager replay with events on the launching
          stream), vs the measured sustained bf16 peak in MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch from the
          committed ncu capture of the SAME plan (profiles/r02_igemm_traffic.json), null when no capture matches
  cpu_baseline  the oracle (fp32 PyTorch restatement of the reference's diffusers path) on this box's host cores, on a
          bounded sample of the same workload (see `sample`); reported, not a target
--impl reference times that CPU path alone (the reference's own code cannot run here: diffusers is not installable).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_COND_XRAY = 15


def balanced_microbatch(total, cap, n_cond=2):
    """the engine's micro-batch size (abi_engine.cu dm_typicality): fewest batches under the cap, equal sizes"""
    cap = max(n_cond, cap // n_cond * n_cond)
    n_mb = -(-total // cap)
    return -(-(-(-total // n_mb)) // n_cond) * n_cond


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"tflops": 1400.0, "source": "fallback (B200_PROFILING.md sustained figure)"}


def load_traffic(key, Bf, aux):
    """DRAM bytes per igemm launch from the committed ncu capture of the same plan (tools/summarize_profiles.py)"""
    tp = os.path.join(ROOT, "profiles", "r02_igemm_traffic.json")
    try:
        d = json.load(open(tp)).get(key)
        if d and int(d.get("Bf", -1)) == Bf and int(d.get("aux", -1)) == aux:
            return d.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_images(torch, n, size, seed):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randint(0, 256, (n, 3, size, size), generator=g, dtype=torch.uint8)
    return imgs.float().div(255.0).mul(2).sub(1)  # load_image(): to_tensor * 2 - 1 (compute.py:126-132)


def make_contexts(torch, n=2):
    g = torch.Generator().manual_seed(5)
    return [torch.randn(77, 768, generator=g) for _ in range(n)]  # [uncond "", cond, ...]; synthetic CLIP states


# ----------------------------------------------------------------------------------------------- workloads
class Typicality:
    """configs 2 and 5: VAE encode -> N (eps,t) draws x n_cond conditions -> raw loss grid + T maps"""

    def __init__(self, cfg):
        self.cfg = cfg
        if cfg == 2:
            self.metric = "typicality samples/sec (512x512, 32 t-steps, cond+uncond)"
            self.images, self.draws_n, self.n_cond, self.img, self.cap = 16, 32, 2, 512, 56
            self.flop_per_unit = 64 * 803.3e9 + 1116.7e9
            self.t_lo, self.t_hi = 100, 700
            self.workload = ("configs[1]: per GPU 16x 512x512 synthetic RGB, 32 (eps,t) draws, cond+uncond = 1024 U-Net forwards "
                             "@64x64 + 16 VAE encodes per step; synthetic seeded SD-1.5 weights (859.5M U-Net, 34.2M VAE enc)")
        else:
            self.metric = "typicality samples/sec (1024x1024, 16 t-steps, 14 conds + uncond)"
            self.images, self.draws_n, self.n_cond, self.img, self.cap = 1, 16, N_COND_XRAY, 1024, 0
            self.flop_per_unit = 240 * 4674.0e9 + 4879.0e9
            self.t_lo, self.t_hi = 0, 1000   # xray/compute.py:103
            self.workload = ("configs[4]: per GPU 1x 1024x1024 synthetic image, 16 (eps,t) draws x (14 conditions + uncond) = 240 U-Net "
                             "forwards @128x128 + 1 VAE encode per step; synthetic seeded SD-1.5 weights")
        self.unit = "samples/s"
        self.lat = self.img // 8

    def setup(self, torch, eng, dev, rank, world):
        self.torch, self.eng, self.dev, self.world = torch, eng, dev, world
        ctxs = make_contexts(torch, self.n_cond)
        for i, c in enumerate(ctxs):
            eng.set_context(i, c)
        self.ctxs = ctxs
        self.slots = list(range(1, self.n_cond)) + [0]  # conditions, then unconditional (compute.py:187-188)
        self.imgs_host = make_images(torch, self.images, self.img, 1000 + rank).pin_memory()
        self.imgs_dev = self.imgs_host.to(dev)
        self.n_total = self.images * world
        self.grid_host = torch.empty(self.images, self.draws_n, self.n_cond, 4, self.lat, self.lat, dtype=torch.float16).pin_memory()
        self.T_host = torch.empty(self.n_total, self.n_cond - 1, self.lat, self.lat, dtype=torch.float32).pin_memory()
        self.pending = []

    def draws(self):
        # D.draws(): re-seeded per image, so every image shares the same N (eps, t) (compute.py:139-141)
        torch, dev = self.torch, self.dev
        torch.manual_seed(42)
        x = torch.empty(1, 4, self.lat, self.lat, device=dev)
        ns, ts = zip(*[(torch.randn_like(x), torch.randint(self.t_lo, self.t_hi, (1,), device=dev)) for _ in range(self.draws_n)])
        return torch.cat(ns), torch.cat(ts).long()

    def _score(self, imgs):
        torch = self.torch
        from diff_mining_b200 import parallel

        post = torch.randn(self.images, 4, self.lat, self.lat, device=self.dev, dtype=torch.float16)
        x0 = self.eng.vae_encode(imgs, post)
        noise, t = self.draws()
        grid, T = self.eng.typicality(x0, noise, t, self.slots, max_forwards=self.cap)
        if self.world > 1:
            h = parallel.gather_tmaps_async(T, self.n_total)   # NCCL all-gather on the process group's stream
            self.pending.append(h)
            return grid, h
        return grid, T

    def step_resident(self):
        return self._score(self.imgs_dev)

    def step_e2e(self):
        x = self.imgs_host.to(self.dev, non_blocking=True)
        grid, T = self._score(x)
        self.grid_host.copy_(grid, non_blocking=True)  # stream-ordered D2H into pinned memory; the timed region ends with
        if self.world > 1:                             # a device synchronize, so everything has landed
            T = T.result()
        self.T_host.copy_(T.reshape(self.T_host.shape), non_blocking=True)
        return self.grid_host, self.T_host

    def finish(self):
        for h in self.pending:
            h.result()
        self.pending = []

    def bytes_per_step(self):
        return self.imgs_host.numel() * 4, self.grid_host.numel() * 2 + self.T_host.numel() * 4

    def plan(self):
        Bf = balanced_microbatch(self.images * self.draws_n * self.n_cond,
                                 self.cap if self.cap else max(8, min(56, 56 * 4096 // (self.lat * self.lat))), self.n_cond)
        return "unet", Bf, self.lat, self.lat, self.n_cond

    def config(self):
        _, Bf, _, _, _ = self.plan()
        return {"workload": self.workload, "images_per_gpu_step": self.images, "mc_samples": self.draws_n, "n_cond": self.n_cond,
                "micro_batch_forwards": Bf,
                "l2": "inputs larger than L2: each micro-batch streams 1.72 GB of weights + >1 GB of activations through a 126 MB L2; no explicit flush"}

    def check(self, out):
        grid = out[0]
        assert self.torch.isfinite(grid.float()).all()

    def cpu_sample(self, torch, sd15, usd, vsd, threads):
        """bounded sample on the host: 1 image -> VAE encode + ONE (eps,t) draw through {c, uncond} (config 2) or through
        one condition (config 5); a full image costs t_vae + draws * n_cond * t_forward"""
        torch.set_num_threads(threads)
        g = torch.Generator().manual_seed(0)
        img = torch.rand(1, 3, self.img, self.img, generator=g) * 2 - 1
        nf = 2 if self.cfg == 2 else 1
        with torch.no_grad():
            t0 = time.perf_counter()
            mean, logvar = sd15.vae_encode_moments(vsd, img)
            x = sd15.vae_sample(mean, logvar, torch.randn(mean.shape, generator=g))
            t1 = time.perf_counter()
            noise = torch.randn(1, 4, self.lat, self.lat, generator=g)
            t = torch.randint(100, 700, (1,), generator=g)
            noisy = sd15.add_noise(x, noise, t).expand(nf, -1, -1, -1)
            pred = sd15.unet_forward(usd, noisy, t.expand(nf), torch.stack([self.ctxs[1], self.ctxs[0]][:nf]))
            loss = (pred - noise) ** 2
            t2 = time.perf_counter()
        assert torch.isfinite(loss).all()
        t_vae, t_fwd = t1 - t0, (t2 - t1) / nf
        v = 1.0 / (t_vae + self.draws_n * self.n_cond * t_fwd)
        sample = (f"1 image {self.img}x{self.img}: VAE encode ({t_vae:.1f} s) + {nf} of {self.draws_n * self.n_cond} U-Net forwards "
                  f"@{self.lat}x{self.lat} ({t_fwd:.1f} s each), fp32 oracle on all host threads; value = 1/(t_vae + {self.draws_n * self.n_cond}*t_forward) [extrapolated]")
        return v, sample


class Geo:
    """config 4: the whole 10k-image data set per step, sharded; one all-gather of the T maps at the end"""

    def __init__(self):
        self.metric = "typicality samples/sec (512x512, 32 t-steps, cond+uncond)"
        self.unit = "samples/s"
        self.total = int(os.environ.get("DM_BENCH_C4_IMAGES", "10000"))
        self.draws_n, self.n_cond, self.img, self.lat, self.cap, self.chunk = 32, 2, 512, 64, 56, 16
        self.n_countries = 10
        self.flop_per_unit = 64 * 803.3e9 + 1116.7e9
        self.workload = (f"configs[3]: {self.total} synthetic 512x512 RGB images in total, sharded i::world, 10 country prompts + uncond, 32 (eps,t) "
                         "draws x {country, uncond} per image = 64 U-Net forwards @64x64 + 1 VAE encode; ONE all-gather of all T maps per step")

    def setup(self, torch, eng, dev, rank, world):
        from diff_mining_b200 import parallel

        self.torch, self.eng, self.dev, self.world, self.rank = torch, eng, dev, world, rank
        ctxs = make_contexts(torch, self.n_countries + 1)
        for i, c in enumerate(ctxs):
            eng.set_context(i, c)
        self.ctxs = ctxs
        self.mine = parallel.shard_indices(self.total, world, rank)
        self.images = len(self.mine)
        # uint8 images, as a loader would hand them over (load_image: to_tensor * 2 - 1 happens on the device)
        g = torch.Generator().manual_seed(4000 + rank)
        pool = torch.randint(0, 256, (64, 3, self.img, self.img), generator=g, dtype=torch.uint8)
        self.imgs_host = pool.pin_memory()       # 64 distinct images, cycled: content does not change the work
        self.imgs_dev = self.imgs_host.to(dev)
        self.grid_host = [torch.empty(self.chunk, self.draws_n, 2, 4, self.lat, self.lat, dtype=torch.float16).pin_memory() for _ in range(2)]
        self.T_host = torch.empty(self.total, self.lat, self.lat, dtype=torch.float32).pin_memory()
        self.copy_stream = torch.cuda.Stream(dev)
        # images of one country are scored together (2 conditions per image: its country and uncond, parallel-dataset/compute.py:153-154)
        self.chunks = []
        for c in range(self.n_countries):
            loc = [k for k, i in enumerate(self.mine) if i % self.n_countries == c]
            self.chunks += [(c, loc[a:a + self.chunk]) for a in range(0, len(loc), self.chunk)]

    def draws(self):
        torch, dev = self.torch, self.dev
        torch.manual_seed(42)
        x = torch.empty(1, 4, self.lat, self.lat, device=dev)
        ns, ts = zip(*[(torch.randn_like(x), torch.randint(100, 700, (1,), device=dev)) for _ in range(self.draws_n)])
        return torch.cat(ns), torch.cat(ts).long()

    def _run(self, e2e):
        torch = self.torch
        from diff_mining_b200 import parallel

        T_local = torch.empty(self.images, self.lat, self.lat, device=self.dev)
        noise, t = self.draws()
        evs = [None, None]
        for ci, (c, loc) in enumerate(self.chunks):
            sel = torch.tensor([k % 64 for k in loc])
            if e2e:
                u8 = self.imgs_host[sel].pin_memory().to(self.dev, non_blocking=True)
            else:
                u8 = self.imgs_dev[sel.to(self.dev)]
            imgs = u8.float().div_(255.0).mul_(2).sub_(1)
            post = torch.randn(len(loc), 4, self.lat, self.lat, device=self.dev, dtype=torch.float16)
            x0 = self.eng.vae_encode(imgs, post)
            grid, T = self.eng.typicality(x0, noise, t, [1 + c, 0], max_forwards=self.cap)
            T_local[torch.tensor(loc, device=self.dev)] = T[:, 0]
            if e2e:   # the per-image .npy payloads leave the device through a double-buffered pinned ring on a side stream
                b = ci & 1
                if evs[b] is not None:
                    evs[b].synchronize()
                done = torch.cuda.Event()
                done.record()
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(done)
                    self.grid_host[b][: len(loc)].copy_(grid, non_blocking=True)
                    grid.record_stream(self.copy_stream)
                    evs[b] = torch.cuda.Event()
                    evs[b].record(self.copy_stream)
        Tall = parallel.gather_tmaps(T_local, self.total) if self.world > 1 else T_local   # the ONE collective of the step
        if e2e:
            torch.cuda.current_stream().wait_stream(self.copy_stream)
            self.T_host.copy_(Tall, non_blocking=True)
            return self.grid_host[0], self.T_host
        return T_local, Tall

    def step_resident(self):
        return self._run(False)

    def step_e2e(self):
        return self._run(True)

    def finish(self):
        pass

    def bytes_per_step(self):
        return self.images * 3 * self.img * self.img, self.images * self.draws_n * 2 * 4 * self.lat * self.lat * 2 + self.total * self.lat * self.lat * 4

    @property
    def units(self):
        return self.images

    def plan(self):
        return "unet", balanced_microbatch(self.chunk * self.draws_n * 2, self.cap, 2), self.lat, self.lat, 2

    def config(self):
        return {"workload": self.workload, "images_total": self.total, "images_this_rank": self.images, "mc_samples": self.draws_n, "n_cond": 2,
                "contexts_resident": self.n_countries + 1, "micro_batch_forwards": self.plan()[1],
                "l2": "inputs larger than L2: each micro-batch streams 1.72 GB of weights + >1 GB of activations through a 126 MB L2; no explicit flush"}

    def check(self, out):
        assert self.torch.isfinite(out[1]).all()

    cpu_sample = Typicality.cpu_sample
    cfg = 2
    images_cfg = None


class Dift:
    """config 3: SDFeaturizer.forward on 8 images x ensemble 8 (64 members per step)"""

    def __init__(self):
        self.metric = "DIFT-161 feature members/sec (512x512, t=161, up_ft_index=1)"
        self.unit = "members/s"
        self.images, self.E, self.img, self.t = 8, 8, 512, 161
        self.lat = 64
        self.flop_per_unit = 438.8e9 + 1116.7e9 / self.E  # one partial forward + 1/8 of the (encode-once) VAE pass
        self.workload = ("configs[2]: per GPU 8x 512x512 synthetic RGB x ensemble 8 = 64 DIFT members per step: 8 VAE encodes (once "
                         "per image), 64 partial U-Net forwards (through up_blocks[1]) @64x64, t=161, ensemble mean -> [8,1280,32,32]")

    def setup(self, torch, eng, dev, rank, world):
        from diff_mining_b200.dift import SDFeaturizer

        self.torch, self.eng, self.dev, self.world = torch, eng, dev, world
        self.ctxs = make_contexts(torch, 2)
        self.f = SDFeaturizer(None, engine=eng, prompt_embeds={"a car": self.ctxs[1]}, device=dev)
        self.imgs_host = make_images(torch, self.images, self.img, 3000 + rank).pin_memory()
        self.imgs_dev = self.imgs_host.to(dev)
        self.feat_host = torch.empty(self.images, 1280, 32, 32, dtype=torch.float32).pin_memory()

    def step_resident(self):
        return (self.f.forward(self.imgs_dev, "a car", t=self.t, up_ft_index=1, ensemble_size=self.E),)

    def step_e2e(self):
        x = self.imgs_host.to(self.dev, non_blocking=True)
        ft = self.f.forward(x, "a car", t=self.t, up_ft_index=1, ensemble_size=self.E)
        self.feat_host.copy_(ft, non_blocking=True)
        return (self.feat_host,)

    def finish(self):
        pass

    def bytes_per_step(self):
        return self.imgs_host.numel() * 4, self.feat_host.numel() * 4

    @property
    def units(self):
        return self.images * self.E

    def plan(self):
        return "dift", 64, self.lat, self.lat, 1

    def config(self):
        return {"workload": self.workload, "images_per_gpu_step": self.images, "ensemble": self.E, "members_per_gpu_step": self.images * self.E,
                "micro_batch_forwards": 64,
                "l2": "inputs larger than L2: one 64-member micro-batch streams 1.54 GB of weights + >1 GB of activations; no explicit flush"}

    def check(self, out):
        assert self.torch.isfinite(out[0]).all()

    def cpu_sample(self, torch, sd15, usd, vsd, threads):
        torch.set_num_threads(threads)
        g = torch.Generator().manual_seed(0)
        img = torch.rand(1, 3, self.img, self.img, generator=g) * 2 - 1
        with torch.no_grad():
            t0 = time.perf_counter()
            mean, logvar = sd15.vae_encode_moments(vsd, img)
            x = sd15.vae_sample(mean, logvar, torch.randn(mean.shape, generator=g))
            t1 = time.perf_counter()
            noise = torch.randn(1, 4, self.lat, self.lat, generator=g)
            tt = torch.full((1,), self.t)
            ft = sd15.unet_forward(usd, sd15.add_noise(x, noise, tt), tt, self.ctxs[1][None], up_ft_index=1)
            t2 = time.perf_counter()
        assert torch.isfinite(ft).all()
        t_vae, t_fwd = t1 - t0, t2 - t1
        # the reference encodes every ensemble member (dift.py:187,220): one member = one encode + one partial forward
        v = 1.0 / (t_vae + t_fwd)
        return v, (f"1 DIFT member 512x512: VAE encode ({t_vae:.1f} s) + partial U-Net forward through up_blocks[1] ({t_fwd:.1f} s), fp32 "
                   "oracle on all host threads, per-member VAE encode as the reference does")


def make_workload(cfg):
    if cfg in (2, 5):
        return Typicality(cfg)
    if cfg == 4:
        return Geo()
    if cfg == 3:
        return Dift()
    raise SystemExit(f"--config must be 2, 3, 4 or 5 (got {cfg})")


# ----------------------------------------------------------------------------------------------- arms
def run_reference(args):
    import torch

    from oracle import sd15

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.config)
    wl.ctxs = make_contexts(torch, 2)
    threads = os.cpu_count() or 1
    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    vsd = sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1)
    for _ in range(max(0, min(args.warmup, 1))):  # one warm-up is enough on a CPU; each costs ~10-30 s
        wl.cpu_sample(torch, sd15, usd, vsd, threads)
    vals, steps_ms, sample = [], [], ""
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, sample = wl.cpu_sample(torch, sd15, usd, vsd, threads)
        steps_ms.append((time.perf_counter() - t0) * 1e3)
        vals.append(v)
    value = statistics.median(vals)
    line = {"metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": statistics.median(steps_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl.workload + " (bounded CPU sample per step)",
                       "note": "reference's diffusers path restated in PyTorch (oracle/sd15.py); diffusers itself is not installable offline"},
            "cpu_baseline": {"value": value, "unit": wl.unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from diff_mining_b200.engine import Engine
    from oracle import sd15  # synthetic weight generator + the cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    vsd = sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1)
    eng = Engine(local)
    eng.load_state_dict(usd, "unet.")
    eng.load_state_dict(vsd, "vae.")
    eng.finalize()
    eng.set_schedule(*sd15.schedule_tables())
    wl = make_workload(args.config)
    wl.setup(torch, eng, dev, rank, world)
    units = getattr(wl, "units", None) or wl.images
    n_total = wl.total if args.config == 4 else units * world   # config 4: fixed total work (strong scaling)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count
        e0.record()
        for _ in range(steps):
            out = fn()
        wl.finish()   # outstanding T-map all-gathers complete inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, eng.launch_count - l0, out

    warm = max(args.warmup, 3)
    if args.config == 4:
        # one step = the whole data set (minutes): warm up on three chunks instead of three full passes
        full, wl.chunks = wl.chunks, wl.chunks[:3]
        for _ in range(warm):
            wl.step_resident()
            wl.step_e2e()
        wl.chunks = full
        warm_note = f"{warm} x 3 chunks of 16 images (a full pass is one timed step)"
    for _ in range(0 if args.config == 4 else warm):
        wl.step_resident()
    wl.finish()
    clocks = ClockSampler(local)
    clocks.start()
    ms_step, launches, out = timed(wl.step_resident, args.steps)
    clk = clocks.stop()
    wl.check(out)
    for _ in range(0 if args.config == 4 else 2):
        wl.step_e2e()
    wl.finish()
    ms_e2e, _, out_e2e = timed(wl.step_e2e, args.steps)
    h2d, d2h = wl.bytes_per_step()

    value = n_total / (ms_step * 1e-3)
    pk = load_peaks()
    cfg = wl.config()
    cfg["parallelism"] = f"dp{world} (image sharding" + (", one async all-gather of T maps per step)" if args.config != 3 else ")")
    if args.config == 4:
        cfg["warmup_note"] = warm_note
    line = {"metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.config == 4 else "weak", "vs_baseline": None, "dtype": "fp16",
            "data": "synthetic", "config": cfg, "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": n_total / (ms_e2e * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "tensor_roofline_frac_whole_step": value / world * wl.flop_per_unit / 1e12 / pk["tflops"]}

    if rank == 0:
        # ---- roofline of the dominant kernel: every igemm launch of one micro-batch of the benched plan, events per launch
        kind, Bf, h, w, aux = wl.plan()
        pr = eng.profile_plan(kind, Bf, h, w, aux, iters=3)
        tot = pr["ms_igemm"] + pr["ms_attn"] + pr["ms_other"]
        ach = pr["flops_igemm"] / (pr["ms_igemm"] * 1e-3) / 1e12
        att = pr["flops_attn"] / (pr["ms_attn"] * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                            "traffic": load_traffic(f"config{2 if args.config == 4 else args.config}", Bf, aux),  # config 4 replays config 2's plan
                            "kernel": "dm::igemm_kernel<BN> (tcgen05 implicit GEMM: all convs + Linears)",
                            "plan": {"kind": kind, "Bf": Bf, "h": h, "w": w, "aux": aux},
                            "peak_source": pk["source"],
                            "share_of_unet_microbatch": pr["ms_igemm"] / tot,
                            "attention": {"achieved": att, "unit": "TFLOP/s", "frac": att / pk["tflops"],
                                          "share_of_unet_microbatch": pr["ms_attn"] / tot},
                            "other_share_of_unet_microbatch": pr["ms_other"] / tot}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, sample = wl.cpu_sample(torch, sd15, usd, vsd, threads)
            line["cpu_baseline"] = {"value": v, "unit": wl.unit, "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
