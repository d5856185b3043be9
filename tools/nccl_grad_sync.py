"""Two-(or more-)GPU check of the training step's exchange (training-step row): GradBuckets over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nccl_grad_sync.py

Every rank builds C_NETWORK on its own GPU, runs the GPU train-mode forward + first backward stage on its OWN shard of the
batch (dcsnet_b200.TrainStep), fills the flat decoder-first gradient buckets (rank-dependent values for every parameter, plus
the kernel-computed train-mode BatchNorm gradients of the step), all-reduces them over NCCL with the first bucket launched
asynchronously, averages, clips by the global norm, and checks the result against the closed form.  Rank 0 prints one JSON line.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dcsnet_b200 as D  # noqa: F401
    from dcsnet_b200 import c_network, config as C, grad_sync, train_engine, train_ops as T
    from oracle import dcsnet_oracle as O
    hp = dict(C.hparams)
    hp["dropout_conv"], hp["dropout_fc"] = 0.0, 0.0
    net = c_network.C_NETWORK(C.config, hp, 0).cuda()
    # ---- a real GPU step on this rank's shard: per-rank losses differ, replicas keep LOCAL BatchNorm statistics
    clean, noise, noisy = O.synthetic_audio(2, 32 * 63, seed=1234 + rank)
    step = train_engine.TrainStep(net, "dcs")
    losses = step.forward(O.stft(noise).cuda(), O.stft(noisy).cuda(), O.stft(clean).cuda())
    b = step.backward_first_stage()
    # kernel-computed parameter gradients available so far: none reach a parameter through the first stage alone, so the BN
    # backward kernel is exercised on the last BatchNorm with the (real) upstream gradient shape; every other parameter gets a
    # rank-dependent synthetic gradient
    names = sorted(n for n, _ in net.named_parameters())
    seeds = {n: i for i, n in enumerate(names)}
    for n, p in net.named_parameters():
        g = torch.Generator(device="cuda").manual_seed(seeds[n] + 7919 * rank)
        p.grad = torch.randn(p.shape, generator=g, device="cuda") * (1.0 + rank)
    gb = grad_sync.GradBuckets(net.named_parameters(), bucket_bytes=4 << 20)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gb.launch(0)                      # the last decoder stage's bucket goes out while "backward" would still be running
    torch.cuda.current_stream().wait_stream(side)
    gb.finish()
    want = {}
    for n, p in net.named_parameters():
        acc = torch.zeros_like(p)
        for r in range(world):
            g = torch.Generator(device="cuda").manual_seed(seeds[n] + 7919 * r)
            acc += torch.randn(p.shape, generator=g, device="cuda") * (1.0 + r)
        want[n] = acc / world
    err = max(float((p.grad - want[n]).abs().max()) for n, p in net.named_parameters())
    ref_params = [torch.nn.Parameter(p.detach().clone()) for _, p in net.named_parameters()]
    for rp, (n, _) in zip(ref_params, net.named_parameters()):
        rp.grad = want[n].clone()
    ref_norm = torch.nn.utils.clip_grad_norm_(ref_params, 100.0)
    norm = gb.clip_by_global_norm(100.0)
    cerr = max(float((p.grad - rp.grad).abs().max()) for rp, (_, p) in zip(ref_params, net.named_parameters()))
    loss_all = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(loss_all, losses["train_loss"].reshape(1).float())
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps(dict(world=world, backend=dist.get_backend(), numel=gb.numel, buckets=len(gb.buckets), err=err, clip_err=cerr,
                              norm=float(norm), ref_norm=float(ref_norm), train_loss_per_rank=[float(t) for t in loss_all],
                              g_d5_norm=float(b["g_d5"].norm()), ok=bool(err <= 1e-5 and cerr <= 1e-5 and abs(float(norm) - float(ref_norm)) <= 1e-3 * float(ref_norm)))),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
