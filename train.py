#!/usr/bin/env python
"""train.py — entry point with the reference's command line (`python train.py [dcs|drs|dc|dr] <gpu index>`, /root/reference/train.py:108-152).

The reference builds the VoiceBank loaders, the network of the variant and a Lightning Trainer (gradient clip 100 by norm, SWA,
ReduceLROnPlateau) and calls `trainer.fit`.  What this shim runs ON THE GPU per step, for the complex variants (dcs / dc):
`network.training_step(batch, idx)` = train-mode `C_NETWORK.forward` (batch-statistic BatchNorm, running-stat update, dropout) +
`calc_loss`, then `network.train_step.backward()` (the whole backward pass as hand-written kernels) and
`network.train_step.optimizer_step()` (NCCL all-reduce of the flat gradient buckets when launched under torchrun, global-norm clip,
Adam-amsgrad) — `dcsnet_b200.train_engine.TrainStep`.  The real variants (dr / drs) exit with a clear message.  Data: seeded synthetic
batches (no VoiceBank data in the image).  SWA / ReduceLROnPlateau / checkpointing are Lightning's (out of scope).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variant", choices=["dcs", "drs", "dc", "dr"])
    ap.add_argument("gpu", type=int, nargs="?", default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--frames", type=int, default=256)
    a = ap.parse_args()
    if a.variant in ("dr", "drs"):
        raise SystemExit("train.py: the real path's (dr / drs) training step is not built; dcs / dc run the whole step on the GPU")
    import torch
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network, config as cfg, ops
    from dcsnet_b200 import synthetic as O   # seeded synthetic utterances (no dataset in the image)
    if not torch.cuda.is_available():
        raise SystemExit("train.py needs a CUDA device (sm_100a); dcsnet_b200 has no CPU fallback")
    torch.cuda.set_device(a.gpu)
    hp = dict(cfg.hparams)
    network = c_network.C_NETWORK(cfg.config, hp, cfg.config.seed).cuda().train()
    network.variant = a.variant
    log = []
    for idx in range(a.steps):
        clean, noise, noisy = O.synthetic_audio(a.batch, 32 * (a.frames - 1), seed=2000 + idx)
        batch = (ops.stft(noise.cuda()), ops.stft(noisy.cuda()), ops.stft(clean.cuda()), [f"synthetic_{idx}"] * a.batch)
        loss = network.training_step(batch, idx)
        network.train_step.backward()
        sumsq = network.train_step.optimizer_step()
        log.append({"step": idx, "loss": float(loss), "grad_norm": float(sumsq.sqrt())})
    torch.cuda.synchronize()
    print(json.dumps({"variant": a.variant, "gpu": a.gpu, "steps": log, "optimizer_step": True,
                      "note": "train-mode forward + losses + whole backward pass + clip + Adam-amsgrad, all hand-written sm_100a kernels"}))


if __name__ == "__main__":
    main()
