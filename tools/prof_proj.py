"""Debug counters (per-role wait cycles) of the last cconv_tc launch inside ops.clstm = a layer-1 LSTM input projection."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dcsnet_b200 as D
from dcsnet_b200 import ops
import bench
B, T = 64, 2000
sd = bench.make_weights()
plan = D.ForwardPlan(D.PackedNet(sd, "cuda", "fp16"), B, T, want_aux=False)
g = torch.Generator().manual_seed(0)
plan.audio_in.copy_(0.1 * torch.randn(B, 32 * (T - 1), generator=g))
plan._enqueue_from_audio(); torch.cuda.synchronize()
x = plan.enc[-1]
Bx, H, W, _, _ = x.shape
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
D._lib.lib().dcs_tc_set_debug_buffer(D._lib.ptr(dbg))
ops.clstm(x.view(Bx, H * W, x.shape[3], 2), plan.lat.view(Bx, H * W, 128, 2), plan.pk.lstm, plan.lstm_ws, use_tc=True)
torch.cuda.synchronize()
D._lib.lib().dcs_tc_set_debug_buffer(None)
d = dbg.view(148, 8).double()
act = d[:, 7] > 0
m = d[act].mean(0)
print(json.dumps({"ctas": int(act.sum()), "kcycles": {"prod_wait_empty": round(float(m[0]) / 1e3, 1), "prod_total": round(float(m[1]) / 1e3, 1),
      "mma_wait_full": round(float(m[2]) / 1e3, 1), "mma_wait_acc": round(float(m[3]) / 1e3, 1), "mma_total": round(float(m[4]) / 1e3, 1),
      "epi_wait": round(float(m[5]) / 1e3, 1), "epi_total": round(float(m[6]) / 1e3, 1), "tiles_per_cta": round(float(m[7]), 1)}}))
