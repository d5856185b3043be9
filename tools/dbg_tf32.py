import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import dcsnet_oracle as O
import dcsnet_b200
from dcsnet_b200 import config as cfg, c_network, train_engine
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 64
hp = dict(cfg.hparams); hp["dropout_conv"] = hp["dropout_fc"] = 0.0
clean, noise, noisy = O.synthetic_audio(B, 32 * (T - 1))
specs = (O.stft(noise).cuda(), O.stft(noisy).cuda(), O.stft(clean).cuda())
res = {}
for mode in ("fp32", "tf32"):
    net = c_network.C_NETWORK(cfg.config, hp, 0).cuda()
    st = train_engine.TrainStep(net, "dcs", mode=mode)
    out = st.forward(*specs); info = st.backward(); torch.cuda.synchronize()
    sv = st.saved
    t = {k: v.clone() for k, v in sv.items() if torch.is_tensor(v) and v.dtype == torch.float32}
    t["raw"] = torch.view_as_real(sv["raw"]).clone()
    for k, v in info.items(): t["info_" + k] = v.clone()
    res[mode] = (t, {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}, float(out["train_loss"]))
a, b = res["fp32"], res["tf32"]
print("loss", a[2], b[2])
rel = lambda x, y: float((x - y).norm() / y.norm().clamp_min(1e-30))
for k in a[0]:
    if a[0][k].shape == b[0][k].shape and a[0][k].numel() > 16: print(f"{k:20s} {rel(b[0][k], a[0][k]):.3e}")
rows = sorted(((rel(b[1][k], a[1][k]), k, float(a[1][k].norm())) for k in a[1]), reverse=True)
order = [f"decoder.{i}." + ("" if i == 6 else "0.") + "conv_tran_r.weight" for i in range(6, -1, -1)] + ["fc.fc_r.weight", "lstm.real_lstm.weight_ih_l1", "lstm.real_lstm.weight_ih_l0"] + [f"encoder.{i}.0.conv_r.weight" for i in range(6, -1, -1)]
for k in order: print(f"{rel(b[1][k], a[1][k]):.3e} {float(a[1][k].norm()):.3e} {k}")
num = sum(float((b[1][k] - a[1][k]).double().pow(2).sum()) for k in a[1]) ** 0.5; den = sum(float(a[1][k].double().pow(2).sum()) for k in a[1]) ** 0.5
print("global", num / den)
