"""Time the training step (TrainStep forward / backward / optimizer) at the BASELINE configs[4] size on one GPU.
usage: python tools/time_train.py [--batch 32] [--frames 2000] [--iters 3] [--mode fp32]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--mode", default="fp32")
    ap.add_argument("--no-opt", action="store_true")
    a = ap.parse_args()
    import dcsnet_b200 as D
    from dcsnet_b200 import c_network, config as cfg, ops, train_engine
    from oracle import dcsnet_oracle as O
    torch.cuda.set_device(0)
    net = c_network.C_NETWORK(cfg.config, dict(cfg.hparams), 0).cuda().train()
    kw = {} if a.mode == "fp32" else {"mode": a.mode}
    step = train_engine.TrainStep(net, "dcs", **kw)
    if not a.no_opt:
        step.init_optimizer()
    clean, noise, noisy = O.synthetic_audio(a.batch, 32 * (a.frames - 1))
    specs = [ops.stft(t.cuda()) for t in (noise, noisy, clean)]
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    rows = []
    for it in range(a.iters + 1):
        e = [ev() for _ in range(4)]
        n0 = D._lib.launch_count()
        e[0].record()
        out = step.forward(*specs)
        e[1].record()
        step.backward()
        e[2].record()
        if not a.no_opt:
            step.optimizer_step()
        e[3].record()
        torch.cuda.synchronize()
        rows.append(dict(fwd_ms=e[0].elapsed_time(e[1]), bwd_ms=e[1].elapsed_time(e[2]), opt_ms=e[2].elapsed_time(e[3]),
                         total_ms=e[0].elapsed_time(e[3]), loss=float(out["train_loss"]), launches=D._lib.launch_count() - n0))
    print(json.dumps(dict(batch=a.batch, frames=a.frames, mode=a.mode, steps=rows, peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30)))


if __name__ == "__main__":
    main()
