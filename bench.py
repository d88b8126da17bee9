#!/usr/bin/env python
"""bench.py -- headline benchmark of the typicality hot path (BASELINE.json: "typicality samples/sec (512x512,
32 t-steps, cond+uncond)").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): per GPU and step, 16 synthetic 512x512 RGB images, each scored with 32 (eps,t)
draws x {c, uncond} = 64 SD-1.5 U-Net forwards at 64x64 latents + 1 VAE encode (52.5 TFLOP / image).
One "sample" = one image scored.  Weak scaling: every rank scores its own 16 images; the only exchange is the
all-gather of the per-image T maps (inside the timed region).

Printed JSON (one line, rank 0):
  value   images/s, inputs resident in HBM (fp32 images on the device), outputs left on the device
  e2e     images/s through the host-facing path: pinned host images -> H2D -> VAE -> MC typicality -> D2H of the raw
          fp16 loss grid [16,32,2,4,64,64] (the reference's .npy payload) and the T maps
  roofline   the dominant kernel (tcgen05 implicit GEMM, all launches of one micro-batch): algorithmic FLOPs / CUDA-event
          time per launch, measured live (eager replay with events on the launching stream), vs the measured
          sustained bf16 peak in MEASURED_PEAKS.json
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

METRIC = "typicality samples/sec (512x512, 32 t-steps, cond+uncond)"
UNIT = "samples/s"
IMAGES_PER_STEP = 16
N_DRAWS = 32
IMG = 512
LAT = IMG // 8
MICRO_BATCH = 56  # cap; the engine balances: 1024 forwards -> 19 micro-batches of 54 / 52 (fills 148 SMs at every level)
FLOP_PER_SAMPLE = 64 * 803.3e9 + 1116.7e9


def balanced_microbatch(total, cap, n_cond=2):
    """the engine's micro-batch size (abi_engine.cu dm_typicality): fewest batches under the cap, equal sizes"""
    n_mb = -(-total // cap)
    return -(-(-(-total // n_mb)) // n_cond) * n_cond


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"tflops": 1400.0, "source": "fallback (B200_PROFILING.md sustained figure)"}


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


def make_inputs(torch, seed):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randint(0, 256, (IMAGES_PER_STEP, 3, IMG, IMG), generator=g, dtype=torch.uint8)
    return imgs.float().div(255.0).mul(2).sub(1)  # load_image(): to_tensor * 2 - 1 (compute.py:126-132)


def make_contexts(torch):
    g = torch.Generator().manual_seed(5)
    return [torch.randn(77, 768, generator=g) for _ in range(2)]  # [uncond "", cond "1975"-like]; synthetic CLIP states


def cpu_reference_sample(torch, sd15, usd, vsd, ctxs, threads):
    """bounded sample of the workload on the host: 1 image 512x512 -> VAE encode + ONE (eps,t) draw x {c, uncond}.
    A full sample costs t_vae + 32 * t_pair."""
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    img = torch.rand(1, 3, IMG, IMG, generator=g) * 2 - 1
    with torch.no_grad():
        t0 = time.perf_counter()
        mean, logvar = sd15.vae_encode_moments(vsd, img)
        x = sd15.vae_sample(mean, logvar, torch.randn(mean.shape, generator=g))
        t1 = time.perf_counter()
        noise = torch.randn(1, 4, LAT, LAT, generator=g)
        t = torch.randint(100, 700, (1,), generator=g)
        noisy = sd15.add_noise(x, noise, t).expand(2, -1, -1, -1)
        pred = sd15.unet_forward(usd, noisy, t.expand(2), torch.stack([ctxs[1], ctxs[0]]))
        loss = (pred - noise) ** 2
        t2 = time.perf_counter()
    assert torch.isfinite(loss).all()
    t_vae, t_pair = t1 - t0, t2 - t1
    return t_vae, t_pair, 1.0 / (t_vae + N_DRAWS * t_pair)


def run_reference(args):
    import torch

    from oracle import sd15

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    vsd = sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1)
    ctxs = make_contexts(torch)
    for _ in range(max(0, min(args.warmup, 1))):  # one warm-up is enough on a CPU; each costs ~10-30 s
        cpu_reference_sample(torch, sd15, usd, vsd, ctxs, threads)
    vals, steps_ms = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        t_vae, t_pair, v = cpu_reference_sample(torch, sd15, usd, vsd, ctxs, threads)
        steps_ms.append((time.perf_counter() - t0) * 1e3)
        vals.append(v)
    value = statistics.median(vals)
    sample = "1 image 512x512: VAE encode + 1 of 32 (eps,t) draws x {c,uncond} (2 U-Net forwards @64x64), fp32 oracle; value = 1/(t_vae + 32*t_pair)"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": statistics.median(steps_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "configs[1]: 16x 512x512, 32 (eps,t) draws, cond+uncond (bounded CPU sample per step)",
                       "note": "reference's diffusers path restated in PyTorch (oracle/sd15.py); diffusers itself is not installable offline"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from diff_mining_b200 import parallel
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
    ctxs = make_contexts(torch)
    eng.set_context(0, ctxs[0])
    eng.set_context(1, ctxs[1])
    slots = [1, 0]  # condition, then unconditional (compute.py:187-188)

    imgs_host = make_inputs(torch, 1000 + rank).pin_memory()
    imgs_dev = imgs_host.to(dev)
    n_total = IMAGES_PER_STEP * world

    def draws():
        # D.draws(): re-seeded per image, so every image shares the same N (eps, t) (compute.py:139-141)
        torch.manual_seed(42)
        x = torch.empty(1, 4, LAT, LAT, device=dev)
        ns, ts = zip(*[(torch.randn_like(x), torch.randint(100, 700, (1,), device=dev)) for _ in range(N_DRAWS)])
        return torch.cat(ns), torch.cat(ts).long()

    def step_resident():
        post = torch.randn(IMAGES_PER_STEP, 4, LAT, LAT, device=dev, dtype=torch.float16)
        x0 = eng.vae_encode(imgs_dev, post)
        noise, t = draws()
        grid, T = eng.typicality(x0, noise, t, slots, max_forwards=MICRO_BATCH)
        Tall = parallel.gather_tmaps(T[:, 0].contiguous(), n_total) if world > 1 else T
        return grid, Tall

    # pinned landing buffers for the step's results (the reference's .npy payload and the T maps)
    grid_host = torch.empty(IMAGES_PER_STEP, N_DRAWS, 2, 4, LAT, LAT, dtype=torch.float16).pin_memory()
    T_host = torch.empty(n_total if world > 1 else IMAGES_PER_STEP, LAT, LAT, dtype=torch.float32).pin_memory()

    def step_e2e():
        x = imgs_host.to(dev, non_blocking=True)
        post = torch.randn(IMAGES_PER_STEP, 4, LAT, LAT, device=dev, dtype=torch.float16)
        x0 = eng.vae_encode(x, post)
        noise, t = draws()
        grid, T = eng.typicality(x0, noise, t, slots, max_forwards=MICRO_BATCH)
        Tall = parallel.gather_tmaps(T[:, 0].contiguous(), n_total) if world > 1 else T[:, 0]
        grid_host.copy_(grid, non_blocking=True)   # stream-ordered D2H into pinned memory; the timed region ends
        T_host.copy_(Tall.reshape(T_host.shape), non_blocking=True)  # with a device synchronize, so both have landed
        return grid_host, T_host

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
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, eng.launch_count - l0, out

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    ms_step, launches, out = timed(step_resident, args.steps)
    clk = clocks.stop()
    grid, _ = out
    assert torch.isfinite(grid.float()).all()
    for _ in range(2):
        step_e2e()
    ms_e2e, _, out_e2e = timed(step_e2e, args.steps)
    h2d = imgs_host.numel() * 4
    d2h = out_e2e[0].numel() * 2 + out_e2e[1].numel() * 4

    value = n_total / (ms_step * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
            "data": "synthetic",
            "config": {"workload": "configs[1]: per GPU 16x 512x512 synthetic RGB, 32 (eps,t) draws, cond+uncond = 1024 U-Net forwards "
                                   "@64x64 + 16 VAE encodes per step; synthetic seeded SD-1.5 weights (859.5M U-Net, 34.2M VAE enc)",
                       "images_per_gpu_step": IMAGES_PER_STEP, "mc_samples": N_DRAWS, "n_cond": 2, "micro_batch_forwards": balanced_microbatch(IMAGES_PER_STEP * N_DRAWS * 2, MICRO_BATCH),
                       "l2": "inputs larger than L2: each micro-batch streams 1.72 GB of weights + >1 GB of activations through a 126 MB L2; no explicit flush",
                       "parallelism": f"dp{world} (image sharding, one all-gather of T maps)"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": n_total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "tensor_roofline_frac_whole_step": value / world * FLOP_PER_SAMPLE / 1e12 / load_peaks()["tflops"]}

    if rank == 0:
        # ---- roofline of the dominant kernel: every igemm launch of one 32-forward micro-batch, events per launch
        pk = load_peaks()
        pr = eng.profile_unet(balanced_microbatch(IMAGES_PER_STEP * N_DRAWS * 2, MICRO_BATCH), LAT, LAT, iters=3)
        n_ig = None
        ach = pr["flops_igemm"] / (pr["ms_igemm"] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_igemm_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                            "traffic": traffic, "kernel": "dm::igemm_kernel<BN> (tcgen05 implicit GEMM: all convs + Linears)",
                            "peak_source": pk["source"],
                            "share_of_unet_microbatch": pr["ms_igemm"] / (pr["ms_igemm"] + pr["ms_attn"] + pr["ms_other"]),
                            "attention": {"achieved": pr["flops_attn"] / (pr["ms_attn"] * 1e-3) / 1e12, "unit": "TFLOP/s",
                                          "frac": pr["flops_attn"] / (pr["ms_attn"] * 1e-3) / 1e12 / pk["tflops"],
                                          "share_of_unet_microbatch": pr["ms_attn"] / (pr["ms_igemm"] + pr["ms_attn"] + pr["ms_other"])},
                            "other_share_of_unet_microbatch": pr["ms_other"] / (pr["ms_igemm"] + pr["ms_attn"] + pr["ms_other"])}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            t_vae, t_pair, v = cpu_reference_sample(torch, sd15, usd, vsd, ctxs, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"1 image 512x512: VAE encode ({t_vae:.1f} s) + 1 of 32 (eps,t) draws x {{c,uncond}} ({t_pair:.1f} s), "
                                              "fp32 oracle on all host threads; value = 1/(t_vae + 32*t_pair)"}
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
