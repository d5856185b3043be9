"""Generate tests/golden/train_step.pt by EXECUTING THE REFERENCE'S training-step function with autograd (build container
only; TEST INFRASTRUCTURE).  Pin for SURVEY 8f rank 2 (the training step): train_engine.TrainStep's
forward / backward kernels are checked against these numbers (tests/test_train_gpu.py).

    python -m oracle.make_golden_train

`network_functions.train_batch_2_loss` (network_functions.py:210-280), imported unmodified behind oracle/stubs, on
`C_NETWORK(config, hparams, seed=0).train()` (batch-statistics ComplexBatchNorm2d, running-stat updates) and
`R_NETWORK(...)` likewise, variants dcs and drs.  Dropout probabilities are set to 0 through hparams (`dropout_conv`,
`dropout_fc`): torch's CPU dropout stream cannot be reproduced by another implementation, everything else is deterministic.
Same runtime substitution as make_golden_eval.py for the cuda-only window in mag_phase_2_wave.  Stored per variant: the three
losses, for every parameter the gradient's L2 norm, |max| and first 8 values, the global gradient norm (the quantity
`gradient_clip_val` acts on), and the BN running statistics after the step.
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_harness as rh, dcsnet_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B, T = 2, 64


def fp(t):
    t = t.detach()
    if t.is_complex():
        t = torch.view_as_real(t)
    t = t.float()
    return dict(norm=float(t.norm()), max_abs=float(t.abs().max()), head=t.reshape(-1)[:8].clone(), shape=tuple(t.shape))


def main():
    mods = rh.load("dcs")
    nf = mods["network_functions"]
    orig = nf.mag_phase_2_wave
    nf.mag_phase_2_wave = rh.mag_phase_2_wave_cpu
    out = dict(B=B, T=T, audio_seed=1234, dropout=0.0)
    try:
        hp = dict(mods["config"].hparams)
        hp["dropout_conv"], hp["dropout_fc"] = 0.0, 0.0
        clean, noise, noisy = O.synthetic_audio(B, 32 * (T - 1))
        batch = (rh.reference_stft(noise), rh.reference_stft(noisy), rh.reference_stft(clean), ["id0", "id1"])
        for variant, dtype in (("dcs", "complex"), ("drs", "real")):
            with rh.argv_variant(variant):
                if dtype == "complex":
                    net = mods["c_network"].C_NETWORK(mods["config"].config, hp, 0)
                else:
                    net = importlib.import_module("r_network").R_NETWORK(mods["config"].Config(), hp, 0)
                net.train()
                noise_loss, speech_loss, train_loss = nf.train_batch_2_loss(net, batch, 0, dtype)
                train_loss.backward()
            grads = {k: fp(p.grad) for k, p in net.named_parameters() if p.grad is not None}
            missing = [k for k, p in net.named_parameters() if p.grad is None]
            total = float(torch.sqrt(sum((torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).float().pow(2).sum()
                                         for p in net.parameters() if p.grad is not None)))
            stats = {k: fp(v) for k, v in net.state_dict().items() if "running_" in k}
            out[variant] = dict(noise_loss=float(noise_loss.detach()), speech_loss=float(speech_loss.detach()), train_loss=float(train_loss.detach()),
                                grads=grads, no_grad=missing, grad_norm=total, running_stats=stats)
            print(variant, out[variant]["train_loss"], "grad_norm", total, len(grads), "grads", len(missing), "without grad")
        # dropout at the reference's own probabilities (config.py:41-42): torch's CPU stream from a fixed seed.  Another
        # implementation cannot reproduce the stream, but a restatement that draws from the same generator in the same
        # order can — this pins WHERE dropout sits and that real and imaginary parts drop independently (c_network.py:195).
        hp2 = dict(mods["config"].hparams)
        with rh.argv_variant("dcs"):
            net = mods["c_network"].C_NETWORK(mods["config"].config, hp2, 0)
            net.train()
            torch.manual_seed(11)
            nl, sl, tl = nf.train_batch_2_loss(net, batch, 0, "complex")
        out["dcs_dropout"] = dict(seed=11, dropout_conv=hp2["dropout_conv"], dropout_fc=hp2["dropout_fc"],
                                  noise_loss=float(nl.detach()), speech_loss=float(sl.detach()), train_loss=float(tl.detach()))
        print("dcs with dropout", out["dcs_dropout"])
    finally:
        nf.mag_phase_2_wave = orig
    path = os.path.join(OUT, "train_step.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
