import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import dcsnet_b200 as D
from oracle import dcsnet_oracle as O, synthetic_weights as SW
torch.cuda.set_device(0)
sd = SW.make_state_dict(0)
_, _, noisy = O.synthetic_audio(2, 32 * 63)
ref = O.enhance_audio(sd, noisy)["clean_audio"]
rel = lambda a: float((a.cpu() - ref).abs().max() / ref.abs().max())
for mode in ("fp32", "fp16"):
    for graph in (False, True):
        for dev in ("cuda:0", "cuda:1"):
            try:
                enh = D.Enhancer(sd, batch=2, n_samples=32 * 63, mode=mode, device=dev, graph=graph)
                out = enh(noisy)
                torch.cuda.synchronize(dev)
                print(mode, "graph" if graph else "eager", dev, "rel", rel(out), "cur", torch.cuda.current_device(), flush=True)
            except Exception as e:
                print(mode, graph, dev, "EXC", repr(e)[:300], flush=True)
