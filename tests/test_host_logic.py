"""CPU: host-side logic — window splitting, rank sharding (world_size-2 gloo), shape contract, loud failure without CUDA."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

import dcsnet_b200 as D
from dcsnet_b200 import pipeline
from conftest import build_product_net

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frames_and_window_contract():
    assert pipeline.frames_for(8160) == 256 and pipeline.frames_for(63968) == 2000
    with pytest.raises(ValueError):
        pipeline.frames_for(64000)   # T = 2001: the reference fails at the skip torch.cat (c_network.py:214)
    with pytest.raises(ValueError):
        pipeline.frames_for(8161)


def test_split_windows_pads_tail_and_keeps_order():
    a = torch.arange(1, 20, dtype=torch.float32)
    w, n = pipeline.split_windows(a, 8)
    assert w.shape == (3, 8) and n == 19
    assert torch.equal(w.reshape(-1)[:19], a) and torch.all(w.reshape(-1)[19:] == 0)
    w, n = pipeline.split_windows(torch.zeros(0), 8)   # empty input -> one all-zero window, length 0
    assert w.shape == (1, 8) and n == 0


@pytest.mark.parametrize("n,world", [(901, 8), (901, 2), (5, 8), (0, 4), (64, 1)])
def test_shard_range_partitions_everything_once(n, world):
    got = []
    for r in range(world):
        a, b = pipeline.shard_range(n, r, world)
        assert 0 <= a <= b <= n
        got += list(range(a, b))
    assert got == list(range(n))


def test_no_cpu_fallback():
    net = build_product_net()
    x = torch.zeros(1, 256, 8, dtype=torch.complex64)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        net.encoder[0][0](torch.zeros(1, 1, 16, 16, dtype=torch.complex64))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            D.Enhancer(net, 1, 8160)
        with pytest.raises(RuntimeError, match="CUDA"):
            D.ForwardPlan(D.PackedNet(net, "cpu", "fp32"), 1, 8)


def test_train_mode_is_refused_loudly():
    net = build_product_net().train()
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 256, 8, dtype=torch.complex64))


def test_state_dict_roundtrip_and_ckpt_keys():
    net = build_product_net("randbn")
    sd = net.state_dict()
    assert len(sd) == 246 and sum(p.numel() for p in net.parameters()) == 2912707
    assert sd["decoder.6.conv_tran_r.weight"].shape == (16, 1, 3, 3)
    assert sd["encoder.0.1.running_mean"].dtype == torch.complex64
    other = build_product_net("default")
    other.load_state_dict(sd)   # Lightning .ckpt['state_dict'] loads the same way (strict)
    for k, v in other.state_dict().items():
        assert torch.equal(v, sd[k]), k


def _gloo_worker(rank, world, port, n_windows, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from dcsnet_b200 import pipeline as P
    a, b = P.shard_range(n_windows, rank, world)
    # stand-in for the per-rank enhancement: each rank tags its windows; the sharded path has no data collective,
    # only the timing reduction (max over ranks) and an optional gather of results on rank 0
    local = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 4)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank == 0:
        q.put((float(t), torch.cat(gathered, 0)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_with_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, cat = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    assert torch.equal(cat[:, 0], torch.arange(7, dtype=torch.float32))


@pytest.mark.parametrize("kind", ["complex", "real"])
def test_load_from_checkpoint_like_test_py(kind, tmp_path):
    """test.py:20-26 / 36-42: `X_NETWORK.load_from_checkpoint(config=, seed=, checkpoint_path=, hparams_file=, map_location=None)`
    on a Lightning-format checkpoint (state_dict + hyper_parameters under the name `hparams`) and an hparams.yaml whose
    function-valued entry is a python-object tag."""
    from dcsnet_b200 import c_network, r_network, config as C
    cls = c_network.C_NETWORK if kind == "complex" else r_network.R_NETWORK
    src = cls(C.config, dict(C.hparams), 3)
    with torch.no_grad():
        for p in src.parameters():
            p.add_(0.01)
    hp = {k: v for k, v in C.hparams.items() if k != "initialisation_distribution"}
    hp["speech_alpha"] = 0.65
    ckpt = tmp_path / "epoch=0-step=289.ckpt"
    torch.save({"state_dict": src.state_dict(), "hyper_parameters": hp, "hparams_name": "hparams", "epoch": 0, "global_step": 289}, ckpt)
    yml = tmp_path / "hparams.yaml"
    yml.write_text("lr: 0.0001\nspeech_alpha: 0.6\nchannels:\n- 1\n- 16\n- 32\n- 64\n- 128\n- 256\n- 256\n- 256\n"
                   "initialisation_distribution: !!python/name:torch.nn.init.xavier_uniform_ ''\n")
    net = cls.load_from_checkpoint(config=C.config, seed=C.config.seed, checkpoint_path=str(ckpt), hparams_file=str(yml), map_location=None)
    net.eval()
    for k, v in net.state_dict().items():
        assert torch.equal(v, src.state_dict()[k]), k
    assert net.hparams["speech_alpha"] == 0.6 and net.hparams["noise_loss_type"] == C.hparams["noise_loss_type"]
    assert net.hparams["initialisation_distribution"] is C.hparams["initialisation_distribution"]
    net2 = cls.load_from_checkpoint(config=C.config, seed=0, checkpoint_path=str(ckpt))
    assert net2.hparams["speech_alpha"] == 0.65
    torch.save({"model": 1}, tmp_path / "bad.ckpt")
    with pytest.raises(KeyError):
        cls.load_from_checkpoint(config=C.config, seed=0, checkpoint_path=str(tmp_path / "bad.ckpt"))


def _grad_sync_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from dcsnet_b200 import c_network, config as C, grad_sync
    net = c_network.C_NETWORK(C.config, dict(C.hparams), 0)
    names = sorted(n for n, _ in net.named_parameters())
    seeds = {n: i for i, n in enumerate(names)}                      # hash() is salted per process: use a stable index
    for n, p in net.named_parameters():
        g = torch.Generator().manual_seed(seeds[n] + 7919 * rank)
        p.grad = torch.randn(p.shape, generator=g) * (1.0 + rank)
    gb = grad_sync.GradBuckets(net.named_parameters(), bucket_bytes=4 << 20)
    first, last = next(iter(gb.slices)), list(gb.slices)[-1]         # decoder-first layout
    gb.launch(0)                                                     # overlapped bucket, the rest reduced in finish()
    gb.finish()
    want = {}
    for n, p in net.named_parameters():
        acc = torch.zeros_like(p)
        for r in range(world):
            g = torch.Generator().manual_seed(seeds[n] + 7919 * r)
            acc += torch.randn(p.shape, generator=g) * (1.0 + r)
        want[n] = acc / world
    err = max(float((p.grad - want[n]).abs().max()) for n, p in net.named_parameters())
    ref_params = [torch.nn.Parameter(p.detach().clone()) for _, p in net.named_parameters()]
    for rp, (n, _) in zip(ref_params, net.named_parameters()):
        rp.grad = want[n].clone()
    ref_norm = torch.nn.utils.clip_grad_norm_(ref_params, 100.0)
    norm = gb.clip_by_global_norm(100.0)
    cerr = max(float((p.grad - rp.grad).abs().max()) for rp, (_, p) in zip(ref_params, net.named_parameters()))
    views = all(p.grad.data_ptr() == gb.buckets[gb.slices[n][0]][gb.slices[n][1]:].data_ptr() for n, p in net.named_parameters())
    if rank == 0:
        q.put(dict(err=err, cerr=cerr, norm=float(norm), ref_norm=float(ref_norm), numel=gb.numel, nb=len(gb.buckets), first=first, last=last, views=views))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_all_reduce_and_clip_with_gloo():
    """Training-step exchange: flat decoder-first buckets, async all-reduce of the first bucket, mean over ranks,
    global-norm clip after the reduce == torch's clip_grad_norm_ on the averaged gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    r = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert r["numel"] == 2912707 and r["nb"] >= 3 and r["views"]
    # last decoder stage first (decoder_attention.12 / 13 are constructed but unused by forward, c_network.py:218), input BN last
    assert r["first"].split(".")[0] in ("decoder", "decoder_attention") and r["last"].startswith("initial_batchnorm.")
    assert r["err"] <= 1e-6 and r["norm"] > 100.0 and abs(r["norm"] - r["ref_norm"]) <= 1e-3 * r["ref_norm"] and r["cerr"] <= 1e-6


def test_hparams_shim_behaves_like_attribute_dict():
    """Lightning's AttributeDict contract the reference relies on: attribute access, AttributeError (not KeyError) for a
    missing name — so hasattr, getattr-with-default, copy.deepcopy and torch.save(model) work."""
    import copy
    import io
    net = build_product_net()
    hp = net.hparams
    assert hp.no_of_layers == hp["no_of_layers"] == 7
    assert not hasattr(hp, "no_such_key") and getattr(hp, "no_such_key", 5) == 5
    twin = copy.deepcopy(net)
    assert twin.hparams == net.hparams and twin.hparams is not net.hparams
    for (k, a), (_, b) in zip(net.state_dict().items(), twin.state_dict().items()):
        assert torch.equal(a, b), k
    buf = io.BytesIO()
    torch.save(net, buf)


def test_non_default_geometry_is_refused():
    """A model whose config / BN eps differ from the geometry the kernel plan is built for must not be computed
    silently with the default tables."""
    import copy
    net = build_product_net()
    bad = copy.deepcopy(net)
    bad.config = copy.copy(net.config)
    bad.config.strideE = [(2, 2)] * 7
    with pytest.raises(NotImplementedError, match="strideE"):
        D.PackedNet(bad, "cpu", "fp32")
    bad2 = copy.deepcopy(net)
    bad2.encoder[0][1].eps = 1e-3
    with pytest.raises(NotImplementedError, match="eps"):
        D.PackedNet(bad2, "cpu", "fp32")


@pytest.mark.parametrize("mode,dtype", [("fp16", torch.float16), ("bf16", torch.bfloat16), ("tc", torch.float16)])
def test_tensor_core_modes_pack_in_their_storage_type(mode, dtype):
    pk = D.PackedNet(build_product_net(), "cpu", mode)
    assert pk.tc and pk.act_dtype == dtype
    assert all(p.w_tc.dtype == dtype for p in pk.enc + pk.dec)
    assert all(s.w_image.dtype == dtype for s in pk.strip.values())
    assert pk.fc.w_tc32 is not None and pk.fc.w_tc32.dtype == torch.float32


@pytest.mark.gpu
def test_ops_run_on_the_tensors_device_not_the_current_one():
    """a plan / tensors on cuda:1 while the current device is cuda:0 must launch on cuda:1's stream."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from dcsnet_b200 import ops
    from oracle import dcsnet_oracle as O, synthetic_weights as SW
    torch.cuda.set_device(0)
    _, _, noisy = O.synthetic_audio(2, 32 * 63)
    want = O.stft(noisy)
    got = ops.stft(noisy.to("cuda:1"))
    assert got.device.index == 1 and float((got.cpu() - want).abs().max()) < 1e-5
    sd = SW.make_state_dict(0)
    enh = D.Enhancer(sd, batch=2, n_samples=32 * 63, mode="fp16", device="cuda:1")
    assert torch.cuda.current_device() == 0
    out = enh(noisy)
    ref = O.enhance_audio(sd, noisy)["clean_audio"]
    assert float((out - ref).abs().max() / ref.abs().max()) <= 2e-3
    assert torch.cuda.current_device() == 0


def _flat_sync_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from dcsnet_b200 import c_network, config as C, grad_sync
    net = c_network.C_NETWORK(C.config, dict(C.hparams), 0)
    gb = grad_sync.GradBuckets(net.named_parameters(), bucket_bytes=4 << 20, flat=True)
    # what the backward kernels do: write through the parameters' .grad views
    for i, (n, p) in enumerate(net.named_parameters()):
        p.grad.fill_(float(i % 7 + 1) * (rank + 1))
    for i in range(len(gb.buckets)):
        gb.launch(i)
    gb.wait()                                                           # SUM stays in the flat buffer (the fused optimizer divides)
    ok = all(bool((p.grad == float(i % 7 + 1) * 3).all()) for i, (n, p) in enumerate(net.named_parameters()))
    contiguous = all(b.data_ptr() == gb.flat.data_ptr() + 4 * sum(x.numel() for x in gb.buckets[:i]) for i, b in enumerate(gb.buckets))
    if rank == 0:
        q.put(dict(ok=ok, contiguous=contiguous, numel=int(gb.flat.numel()), nb=len(gb.buckets)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_bucket_exchange_with_gloo():
    """The training step's exchange as TrainStep.optimizer_step drives it: ONE flat fp32 gradient buffer whose 4 MB
    slices are the all-reduce buckets and whose per-parameter views are the .grad tensors the backward kernels write; every bucket
    launched asynchronously, wait() leaves the cross-rank SUM in place for the fused clip + Adam kernel (world_size 2, gloo)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_flat_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    r = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert r["ok"] and r["contiguous"] and r["numel"] == 2912707 and r["nb"] == 3


def test_product_side_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py (synthetic inputs outside the timed region, the
    cpu_baseline leg, the reference arm) may import it — not the package, not the train.py / test.py entry points."""
    import glob
    import re
    files = glob.glob(os.path.join(ROOT, "dcs-net_b200", "*.py")) + [os.path.join(ROOT, f) for f in ("train.py", "test.py", "dcsnet_b200.py")]
    assert len(files) > 10
    for f in files:
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
