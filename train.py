#!/usr/bin/env python
"""train.py — entry point with the reference's command line (`python train.py [dcs|drs|dc|dr] <gpu index>`, /root/reference/train.py:108-152).

The reference builds the VoiceBank loaders, the network of the variant and a Lightning Trainer (gradient clip 100 by norm, SWA,
ReduceLROnPlateau) and calls `trainer.fit`.  What this shim runs ON THE GPU per step, for the complex variants (dcs / dc):
`network.training_step(batch, idx)` = train-mode `C_NETWORK.forward` (batch-statistic BatchNorm, running-stat update) + `calc_loss`,
then the first backward stage (loss -> iSTFT adjoint -> mask-tail adjoint -> decoder[6] dgrad), `dcsnet_b200.train_engine.TrainStep`.
The rest of the backward pass and the Adam-amsgrad update are NOT built yet, so parameters are not updated:
the script reports the per-step losses and says so (`"optimizer_step": false`) instead of pretending to train.  The real variants
(dr / drs) exit with a clear message.  Data: seeded synthetic batches (no VoiceBank data in the image); dropout is set to 0
(train-mode dropout kernels are not built).
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
        raise SystemExit("train.py: the real path's training step is not built (SURVEY 8f rank 2); dcs / dc run the GPU slice")
    import torch
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network, config as cfg, ops
    from oracle import dcsnet_oracle as O   # synthetic audio generator only
    if not torch.cuda.is_available():
        raise SystemExit("train.py needs a CUDA device (sm_100a); dcsnet_b200 has no CPU fallback")
    torch.cuda.set_device(a.gpu)
    hp = dict(cfg.hparams)
    hp["dropout_conv"], hp["dropout_fc"] = 0.0, 0.0
    network = c_network.C_NETWORK(cfg.config, hp, cfg.config.seed).cuda().train()
    network.variant = a.variant
    network.configure_optimizers()          # the reference's Adam-amsgrad + ReduceLROnPlateau objects (not stepped, see above)
    log = []
    for idx in range(a.steps):
        clean, noise, noisy = O.synthetic_audio(a.batch, 32 * (a.frames - 1), seed=2000 + idx)
        batch = (ops.stft(noise.cuda()), ops.stft(noisy.cuda()), ops.stft(clean.cuda()), [f"synthetic_{idx}"] * a.batch)
        loss = network.training_step(batch, idx)
        grads = network.train_step.backward_first_stage()
        log.append({"step": idx, "loss": float(loss), "d_raw_norm": float(torch.view_as_real(grads["d_raw"]).norm()),
                    "g_d5_norm": float(grads["g_d5"].norm())})
    torch.cuda.synchronize()
    print(json.dumps({"variant": a.variant, "gpu": a.gpu, "steps": log, "optimizer_step": False,
                      "note": "train-mode forward + losses + first backward stage on the GPU; remaining backward kernels and the optimizer are not built"}))


if __name__ == "__main__":
    main()
