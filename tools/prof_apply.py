"""One dec_att5-sized (C = 8, 128 x 1000, batch 64) attention pass, for ncu source-level profiling of spat_stats / spat_apply."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dcsnet_b200 as D
from dcsnet_b200 import ops, packing
from oracle import synthetic_weights as SW
C, H, W, B = int(os.environ.get("PA_C", 8)), int(os.environ.get("PA_H", 128)), int(os.environ.get("PA_W", 1000)), 64
sd = SW.make_state_dict(0)
i = {128: 0, 64: 3, 32: 4, 16: 5, 8: 6}[C]
ca = packing.pack_channel_attention(sd, f"skip_attention.{2 * i}.", "cuda")
w7 = packing.pack_spatial_attention(sd, f"skip_attention.{2 * i + 1}.", "cuda")
x = torch.randn(B, H, W, C, 2, device="cuda").bfloat16()
y = torch.empty_like(x)
sums = torch.zeros(B, C, 2, device="cuda")
ops.chan_pool(x, sums)
gate = torch.empty(B, C, 2, device="cuda")
stats = torch.empty(B, H * W, 4, device="cuda")
for _ in range(2):
    ops.spat_stats(x, None, stats, sums=sums, ca=ca, gate_out=gate)
    ops.spat_apply(x, gate, stats, w7, y)
torch.cuda.synchronize()
print("ok")
