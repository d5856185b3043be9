"""Two eager (un-graphed) passes of the tensor-core plan (PROF_VARIANT dcs | dc | dr | drs) at the BASELINE size, for ncu (-s skips
the first pass)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dcsnet_b200 as D
import bench
B, T = int(os.environ.get("PROF_B", 64)), int(os.environ.get("PROF_T", 2000))
variant = os.environ.get("PROF_VARIANT", "dcs")
sd = bench.make_weights(variant)
if variant in ("dr", "drs"):
    plan = D.RealForwardPlan(D.PackedRealNet(sd, "cuda", os.environ.get("PROF_MODE", "fp16")), B, T, variant=variant, want_aux=False)
else:
    plan = D.ForwardPlan(D.PackedNet(sd, "cuda", os.environ.get("PROF_MODE", "fp16")), B, T, variant=variant, want_aux=False)
g = torch.Generator().manual_seed(0)
plan.audio_in.copy_(0.1 * torch.randn(B, 32 * (T - 1), generator=g))
for _ in range(2):
    plan._enqueue_from_audio()
torch.cuda.synchronize()
print("launches", D._lib.launch_count(), "per pass (library kernels; torch fill_ kernels are extra)")
