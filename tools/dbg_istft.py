import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import dcsnet_b200 as D
from dcsnet_b200 import ops
from oracle import dcsnet_oracle as O
for T in (40, 256):
    g = torch.Generator().manual_seed(3)
    spec = torch.complex(torch.randn(1,256,T,generator=g), torch.randn(1,256,T,generator=g))*0.3
    ref = O.spec_to_wave(spec)[0].numpy()
    got = ops.istft(spec.cuda(), atan2_eps=1e-6)[0].cpu().numpy()
    err = np.abs(got-ref)
    print("T", T, "max err", err.max(), "ref max", np.abs(ref).max())
    seg = [float(err[i:i+256].max()) for i in range(0, len(ref), 256)]
    print("seg err", ["%.3f" % s for s in seg[:24]])
    print("ratio", (got[1000:1010]/ref[1000:1010]))
    print("got", got[:8], "ref", ref[:8])
    # correlation with shifted versions
    for sh in (-32,-1,0,1,32):
        a = got[600:3000]; b = ref[600+sh:3000+sh] if 600+sh>=0 else None
        if b is not None and len(b)==len(a): print("shift", sh, float(np.corrcoef(a,b)[0,1]))
    if T==40:
        # even/odd sample check
        print("even err", err[0::2].max(), "odd err", err[1::2].max())
