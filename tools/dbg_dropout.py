import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import dcsnet_oracle as O, train_oracle as TO
import dcsnet_b200
from dcsnet_b200 import config as cfg, c_network, train_engine, train_ops as T
pc, pf = float(sys.argv[1]), float(sys.argv[2])
B, Tn = 2, 64
hp = dict(cfg.hparams); hp["dropout_conv"], hp["dropout_fc"] = pc, pf
net = c_network.C_NETWORK(cfg.config, hp, 0).cuda()
clean, noise, noisy = O.synthetic_audio(B, 32 * (Tn - 1))
specs = (O.stft(noise).cuda(), O.stft(noisy).cuda(), O.stft(clean).cuda())
step = train_engine.TrainStep(net, "dcs", seed=int(sys.argv[3]) if len(sys.argv) > 3 else 0)
sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
params = {k for k, _ in net.named_parameters()}
out = step.forward(*specs); step.backward(); torch.cuda.synchronize()
sv = step.saved
shapes = [sv["enc"][i + 1].shape for i in range(7)] + [sv["dec_in"][0].shape] + [sv["dec_in"][i + 1].shape for i in range(6)] + [(B, 256, Tn, 1, 2)]
tags = [f"enc{i}" for i in range(7)] + ["fc"] + [f"dec{i}" for i in range(7)]
masks = []
for tag, shp in zip(tags, shapes):
    rec = sv.get("drop_" + tag)
    if rec is None:
        masks.append(None); continue
    m = T.dropout(torch.ones(tuple(shp), device="cuda"), rec[0], step.seed + 1000003 * step.steps_done, rec[1]).cpu()
    masks.append(m.reshape(m.shape[0], -1, m.shape[3], 2) if tag == "fc" else m.permute(0, 3, 1, 2, 4).contiguous())
want = TO.train_step(sd, *[s.cpu() for s in specs], params, "dcs", drop=TO.dropout_from_masks(masks))
print("loss", float(out["train_loss"]), want["train_loss"])
total = sum(float(v.double().pow(2).sum()) for v in want["grads"].values()) ** 0.5
rows = []
for k, gw in want["grads"].items():
    gr = dict(net.named_parameters())[k].grad.detach().cpu()
    rows.append((float((gr - gw).abs().max()) / max(float(gw.abs().max()), 1e-5 * total), float((gr - gw).norm() / gw.norm().clamp_min(1e-5 * total)), k))
rows.sort(reverse=True)
for r in rows[:int(os.environ.get("ROWS", "6"))]: print(f"{r[0]:.3e} {r[1]:.3e} {r[2]}")

# ---- intermediate gradients: oracle taps with retain_grad vs the GPU backward's tensors
if os.environ.get("TAPS"):
    eps = O.HPARAMS["atan2_eps"]
    live = {k: (v.detach().clone().requires_grad_(True) if k in params else v) for k, v in sd.items()}
    taps = {}
    class TapDict(dict):
        def __setitem__(self, k, v):
            if v.requires_grad: v.retain_grad()
            super().__setitem__(k, v)
    taps = TapDict()
    mo = O.c_network_forward(live, specs[1].cpu(), O.HPARAMS, taps=taps, explicit_lstm=True, bn=TO.cbn_train({}), drop=TO.dropout_from_masks(masks))
    mo = mo[None] if mo.dim() == 2 else mo
    mask = O.bound_crm(mo, eps)
    prod = TO._mul(specs[1].cpu(), mask)
    wave = lambda s: O.spec_to_wave(s, eps)
    total_l = (1 - 0.7 * (-O.si_snr(wave(specs[0].cpu()), wave(prod)))) + 0.7 * (-O.si_snr(wave(specs[2].cpu()), wave(specs[1].cpu() - prod)))
    total_l.backward()
    step2 = train_engine.TrainStep(net, "dcs", seed=step.seed)
    step2.forward(*specs)
    # re-run the backward capturing stage gradients
    import dcsnet_b200.train_ops as TT
    cap = {}
    orig_att = TT.attention_bwd
    def att(x, dy, *a, **k):
        cap.setdefault("att_dy", []).append(dy.clone())
        return orig_att(x, dy, *a, **k)
    TT.attention_bwd = att
    step2.backward(); torch.cuda.synchronize()
    nchw = lambda t: torch.view_as_complex(t.detach().float().cpu().contiguous()).permute(0, 3, 1, 2)
    # attention_bwd call order: skip6, then for i = 5..0: dec att i, skip att i
    names = ["skip6"] + sum([[f"dec{i}", f"skip{i}"] for i in range(5, -1, -1)], [])
    for nm, dyv in zip(names, cap["att_dy"]):
        w = taps[nm].grad
        gpu = nchw(dyv)
        if nm.startswith("dec"):
            # GPU dy is the gradient AFTER the dropout adjoint; the oracle's tap 'dec{i}' is the attended output BEFORE dropout: same point
            pass
        diff = (gpu - w).abs()
        big = diff > 1e-3 * w.abs().max()
        pix = big.any(dim=1)                              # (B, H, W): pixels with any channel off
        print(f"{nm:8s} rel {float(diff.max() / w.abs().max()):.3e}  elements off {int(big.sum())} of {big.numel()}  pixels off {int(pix.sum())} of {pix.numel()}",
              pix.nonzero()[:4].tolist())
