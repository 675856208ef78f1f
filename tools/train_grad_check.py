"""Gradient accuracy of the Denoise training step at the real TikTok shape (B = 1024, I = 6710, H = 1024): the per-op
path (Denoise.forward + autograd.LinearTN) and the fused step (train_step.DenoiseLossFn), in both precisions, against a
float64 torch restatement of reference Model.py:183-220,385-428 on the same inputs.  Prints per parameter the relative
L2 error, the worst element error relative to the tensor's RMS, and the fraction of gradient elements whose SIGN
differs (Adam's first steps move every weight by -lr * sign(g), so this is what decides the early trajectory).
    python tools/train_grad_check.py [out.json]"""
import json
import math
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from diffmm_b200.Conf import load_config  # noqa: E402
from diffmm_b200.Model import Denoise, GaussianDiffusion  # noqa: E402

ROOT = os.path.abspath(".")
dev = torch.device("cuda:0")
NAMES = ["emb_layer.weight", "emb_layer.bias", "in_layers.0.weight", "in_layers.0.bias", "out_layers.0.weight",
         "out_layers.0.bias", "gate_layer.weight", "gate_layer.bias"]


def truth(gd, den, x0, t, noise, feat, i_embs, sim_weight):
    """float64 restatement; returns loss and gradients in NAMES order."""
    p = {n: v.detach().double().requires_grad_(True) for n, v in den.named_parameters()}
    ta, tb = gd._tables_f32(dev)
    x_t = ta[t].double()[:, None] * x0.double() + tb[t].double()[:, None] * noise.double()
    d = den.time_emb_dim
    half = d // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=dev) / half)
    temp = t[:, None].float() * freqs[None]
    te = torch.cat([torch.cos(temp), torch.sin(temp)], -1).double()
    temb = te @ p["emb_layer.weight"].t() + p["emb_layer.bias"]
    Fd = feat.double()
    proj = x_t @ Fd
    gate = torch.sigmoid(proj @ p["gate_layer.weight"].t() + p["gate_layer.bias"])
    x_t = x_t + (proj * gate) @ Fd.t()
    h = torch.tanh(torch.cat([x_t, temb], -1) @ p["in_layers.0.weight"].t() + p["in_layers.0.bias"])
    out = h @ p["out_layers.0.weight"].t() + p["out_layers.0.bias"]
    mse = ((out - x0.double()) ** 2).mean(-1)
    w = gd.SNR(torch.clamp(t - 1, min=0)) - gd.SNR(t)
    w = torch.where(t == 0, 1.0, w)
    um = out @ Fd
    ui = x0.double() @ i_embs.double()
    sim = 1 - F.cosine_similarity(um, ui, dim=-1)
    loss = (w * mse + sim_weight * sim).mean()
    loss.backward()
    return float(loss), [p[n].grad for n in NAMES]


def ours(gd, den, x0, t, noise, feat, i_embs, fused):
    os.environ["DIFFMM_FUSED_TRAIN"] = "1" if fused else "0"
    den.zero_grad(set_to_none=True)
    loss = gd.training_losses(den, x0, i_embs, feat, timesteps=t, noise=noise).mean()
    loss.backward()
    g = dict(den.named_parameters())
    return float(loss), [g[n].grad.double() for n in NAMES]


def main():
    out = {}
    torch.manual_seed(0)
    cfg0 = load_config(os.path.join(ROOT, "conf", "tiktok.toml"))
    I, B, H = 6710, cfg0.train.batch, 1024
    cfg0.data.user_num, cfg0.data.item_num = 9308, I
    g = torch.Generator(device="cpu").manual_seed(1)
    x0 = torch.zeros(B, I)
    deg = torch.randint(3, 15, (B,), generator=g)
    for b in range(B):
        x0[b, torch.randperm(I, generator=g)[:deg[b]]] = 1.0
    x0 = x0.to(dev)
    noise = torch.randn(B, I, generator=g).to(dev)
    t = torch.randint(0, cfg0.hyper.steps, (B,), generator=g).to(dev)
    feat = F.leaky_relu(torch.randn(I, 64, generator=g) * 0.3).to(dev)            # projected modality features
    i_embs = (torch.rand(I, 64, generator=g) * 2 - 1).mul_(math.sqrt(6 / (I + 64))).to(dev)   # xavier_uniform
    for trained in (0, 1):
        for prec in ("bf16x3", "bf16"):
            cfg = load_config(os.path.join(ROOT, "conf", "tiktok.toml"))
            cfg.data.user_num, cfg.data.item_num = 9308, I
            cfg.base.precision = prec
            torch.manual_seed(0)
            gd = GaussianDiffusion(cfg).to(dev)
            den = Denoise([I, H], [H, I], cfg).to(dev)
            if trained:
                # a few Adam steps on the per-op path: saturates some tanh units, grows the weights off their init
                opt = torch.optim.Adam(den.parameters(), lr=cfg.train.lr)
                os.environ["DIFFMM_FUSED_TRAIN"] = "0"
                for s in range(8):
                    opt.zero_grad()
                    gd.training_losses(den, x0, i_embs, feat, timesteps=(t + s) % cfg.hyper.steps,
                                       noise=torch.roll(noise, s + 1, 0)).mean().backward()
                    opt.step()
            l0, g0 = truth(gd, den, x0, t, noise, feat, i_embs, cfg.hyper.sim_weight)
            for fused in (0, 1):
                l1, g1 = ours(gd, den, x0, t, noise, feat, i_embs, fused)
                key = f"{'trained' if trained else 'init'} {prec} {'fused' if fused else 'per-op'}"
                rows = {}
                for n, a, b in zip(NAMES, g1, g0):
                    rel = float((a - b).norm() / b.norm())
                    worst = float((a - b).abs().max() / b.pow(2).mean().sqrt())
                    sign = float(((a * b) < 0).double().mean())
                    bias = float((a - b).sum() / b.abs().sum())
                    rows[n] = dict(rel_l2=rel, worst_over_rms=worst, sign_flips=sign, signed_bias=bias)
                out[key] = dict(loss=l1, loss_truth=l0, grads=rows)
                print(f"== {key}: loss {l1:.8f} (float64 {l0:.8f})")
                for n, r in rows.items():
                    print(f"   {n:22s} rel_l2 {r['rel_l2']:.3e}  worst/rms {r['worst_over_rms']:.3e}  sign flips {r['sign_flips']:.4%}"
                          f"  bias {r['signed_bias']:+.2e}", flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
