"""R_NETWORK.forward (real path, fp32 kernel sequence) at the BASELINE batch: ms per forward, for DESIGN.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dcsnet_b200 as D
from dcsnet_b200 import r_network, config as C
B, T = int(os.environ.get("PROF_B", 64)), int(os.environ.get("PROF_T", 2000))
net = r_network.R_NETWORK(C.Config(), dict(C.hparams), 0).cuda().eval()
x = torch.rand(B, 256, T, device="cuda")
for _ in range(2):
    net(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = D._lib.launch_count()
e0.record()
for _ in range(3):
    net(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print({"rnet_fp32_ms": round(ms, 2), "B": B, "T": T, "s_audio_per_s": round(B * 32 * (T - 1) / 16000 / (ms / 1e3)), "launches": (D._lib.launch_count() - n0) // 3})
