"""End-to-end parity exploration on the GPU box: engine vs the fp32 oracle (gold) and vs the fp16-autocast
oracle (reference precision), per layer and at the outputs.  Debug helper; the pytest versions are in tests/."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sd15  # noqa: E402
from diff_mining_b200.engine import Engine  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item(), ((a - b).norm() / (b.norm() + 1e-12)).item()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    t0 = time.time()
    usd = sd15.make_synthetic_weights(sd15.unet_param_shapes(), seed=0)
    vsd = sd15.make_synthetic_weights(sd15.vae_encoder_param_shapes(), seed=1)
    print(f"weights generated {time.time()-t0:.1f}s", flush=True)
    eng = Engine(0)
    eng.load_state_dict(usd, "unet.")
    eng.load_state_dict(vsd, "vae.")
    t1 = time.time()
    eng.finalize()
    print(f"engine finalized {time.time()-t1:.1f}s", flush=True)
    a, b = sd15.schedule_tables()
    eng.set_schedule(a, b)
    g = torch.Generator().manual_seed(5)
    ctxs = [torch.randn(77, 768, generator=g) for _ in range(3)]
    for i, c in enumerate(ctxs):
        eng.set_context(i, c)
    usd_g = {k: v.to(dev) for k, v in usd.items()}
    usd_h = {k: v.half() for k, v in usd_g.items()}
    vsd_g = {k: v.to(dev) for k, v in vsd.items()}
    vsd_h = {k: v.half() for k, v in vsd_g.items()}
    ok = True

    if which in ("unet", "all"):
        for (Bf, h, w) in [(2, 32, 32), (3, 24, 40), (2, 33, 47)]:
            x = torch.randn(Bf, 4, h, w, generator=g)
            t = torch.randint(100, 700, (Bf,), generator=g)
            slots = [i % 3 for i in range(Bf)]
            ctx = torch.stack([ctxs[s] for s in slots]).to(dev)
            taps_gold, taps_ac = {}, {}
            with torch.no_grad():
                gold = sd15.unet_forward(usd_g, x.to(dev), t.to(dev), ctx, taps=taps_gold)
                ac = sd15.unet_forward(usd_h, x.to(dev), t.to(dev), ctx, autocast=True, taps=taps_ac)
            eng.debug_keep(True)
            out = eng.unet_eps(x, t, slots)
            torch.cuda.synchronize()
            print(f"--- unet Bf={Bf} {h}x{w}: per-layer max-rel error vs fp32 gold   [engine | autocast-oracle]")
            for name in taps_gold:
                try:
                    e = eng.debug_fetch(name)
                except RuntimeError as ex:
                    print(f"  {name:45s} (no tap: {ex})")
                    continue
                if name == "conv_out":
                    e = e[:, :4]
                r_e = rel(e, taps_gold[name])
                r_a = rel(taps_ac[name], taps_gold[name])
                flag = "" if r_e[0] < max(4 * r_a[0], 2e-3) else "   <<<<<<"
                print(f"  {name:45s} {r_e[0]:.3e} ({r_e[1]:.3e}) | {r_a[0]:.3e} ({r_a[1]:.3e}){flag}")
            eng.debug_keep(False)
            out2 = eng.unet_eps(x, t, slots)
            out3 = eng.unet_eps(x, t, slots)  # graph replay
            torch.cuda.synchronize()
            r_e, r_a = rel(out, gold), rel(ac, gold)
            same = torch.equal(out, out2) and torch.equal(out2, out3)
            print(f"UNET Bf={Bf} {h}x{w}: engine-vs-gold {r_e[0]:.3e} ({r_e[1]:.3e}) autocast-vs-gold {r_a[0]:.3e} ({r_a[1]:.3e}) "
                  f"engine-vs-autocast {rel(out, ac)[0]:.3e}; keep/eager/graph identical: {same}", flush=True)
            ok &= r_e[0] < max(3 * r_a[0], 3e-3) and same

    if which in ("typ", "all"):
        Bi, N, h, w = 2, 4, 32, 32
        x0 = torch.randn(Bi, 4, h, w, generator=g)
        torch.manual_seed(42)
        noise = torch.randn(N, 4, h, w)
        t = torch.randint(100, 700, (N,))
        grid, T = eng.typicality(x0, noise, t, [1, 0], max_forwards=8)
        torch.cuda.synchronize()
        # oracle: the reference loop (compute.py:134-160) in fp32
        ref = torch.empty(Bi, N, 2, 4, h, w)
        with torch.no_grad():
            for i in range(Bi):
                for ci, slot in enumerate([1, 0]):
                    noisy = sd15.add_noise(x0[i:i + 1].expand(N, -1, -1, -1), noise, t)
                    pred = sd15.unet_forward(usd_g, noisy.to(dev), t.to(dev), ctxs[slot].to(dev)[None].expand(N, -1, -1))
                    ref[i, :, ci] = ((pred.cpu() - noise) ** 2)
        r = rel(grid.cpu(), ref)
        Tref = (ref[:, :, 1].mean(2) - ref[:, :, 0].mean(2)).mean(1)
        rT = rel(T.cpu()[:, 0], Tref)
        print(f"TYPICALITY grid-vs-gold {r[0]:.3e} ({r[1]:.3e})  T-vs-gold {rT[0]:.3e} ({rT[1]:.3e}) |T|max {Tref.abs().max():.3e}", flush=True)
        loss, _ = eng.unet_rows(x0[:1], noise, t, [0] * (2 * N), list(range(N)) * 2, [1] * N + [0] * N)
        print("compute_loss rows vs grid:", rel(loss.cpu().view(2, N, 4, h, w).transpose(0, 1), grid[0].float().cpu())[0])
        ok &= r[0] < 2e-2

    if which in ("vae", "all"):
        for (B, H, W) in [(2, 64, 64), (1, 256, 256), (1, 192, 320)]:
            img = (torch.rand(B, 3, H, W, generator=g) * 2 - 1)
            epsd = torch.randn(B, 4, H // 8, W // 8, generator=g).half().float()
            with torch.no_grad():
                m_g, lv_g = sd15.vae_encode_moments(vsd_g, img.to(dev))
                m_a, lv_a = sd15.vae_encode_moments(vsd_h, img.to(dev), autocast=True)
            z, m, lv = eng.vae_encode(img, epsd, return_moments=True)
            torch.cuda.synchronize()
            zg = sd15.vae_sample(m_g, lv_g, epsd.to(dev))
            print(f"VAE B={B} {H}x{W}: mean eng {rel(m, m_g)[0]:.3e} ac {rel(m_a, m_g)[0]:.3e} | logvar eng {rel(lv, lv_g)[0]:.3e} "
                  f"ac {rel(lv_a, lv_g)[0]:.3e} | z eng {rel(z, zg)[0]:.3e}", flush=True)
            ok &= rel(m, m_g)[0] < max(3 * rel(m_a, m_g)[0], 3e-3)

    if which in ("dift", "all"):
        B, E, h, w = 1, 4, 32, 32
        lat = torch.randn(B * E, 4, h, w, generator=g)
        nz = torch.randn(B * E, 4, h, w, generator=g)
        t = 161
        f = eng.dift(lat, nz, t, 2, E, up_ft_index=1)
        torch.cuda.synchronize()
        with torch.no_grad():
            noisy = sd15.add_noise(lat, nz, torch.full((B * E,), t))
            fg = sd15.unet_forward(usd_g, noisy.to(dev), torch.full((B * E,), t, device=dev), ctxs[2].to(dev)[None].expand(B * E, -1, -1),
                                   up_ft_index=1).view(B, E, 1280, h // 2, w // 2).mean(1)
            fa = sd15.unet_forward(usd_h, noisy.to(dev), torch.full((B * E,), t, device=dev), ctxs[2].to(dev)[None].expand(B * E, -1, -1),
                                   up_ft_index=1, autocast=True).float().view(B, E, 1280, h // 2, w // 2).mean(1)
        print(f"DIFT: eng-vs-gold {rel(f, fg)[0]:.3e} ({rel(f, fg)[1]:.3e}) autocast-vs-gold {rel(fa, fg)[0]:.3e}", flush=True)
        ok &= rel(f, fg)[0] < max(3 * rel(fa, fg)[0], 3e-3)

    if which in ("perf", "all"):
        for (Bf, h, w) in [(8, 64, 64), (16, 64, 64), (32, 64, 64)]:
            x = torch.randn(Bf, 4, h, w)
            t = torch.randint(100, 700, (Bf,))
            slots = [i % 2 for i in range(Bf)]
            for _ in range(3):
                eng.unet_eps(x, t, slots)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            xd, td = x.to(dev), t.to(dev)
            e0.record()
            for _ in range(5):
                eng.unet_eps(xd, td, slots)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"PERF unet Bf={Bf} {h}x{w}: {ms:.2f} ms/microbatch  {ms/Bf:.3f} ms/forward  {803.3*Bf/ms:.1f} TFLOP/s", flush=True)
            pr = eng.profile_unet(Bf, h, w, 3)
            print("   profile:", {k: round(v, 3) if v < 1e6 else f"{v:.3e}" for k, v in pr.items()},
                  f"igemm {pr['flops_igemm']/pr['ms_igemm']/1e9:.0f} TF/s attn {pr['flops_attn']/pr['ms_attn']/1e9:.0f} TF/s", flush=True)
    print("E2E", "OK" if ok else "FAILED", f"{time.time()-t0:.1f}s")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
