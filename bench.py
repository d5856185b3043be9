#!/usr/bin/env python
"""bench.py — seconds of 16 kHz audio enhanced per second by the DCS-Net forward hot path
(STFT -> C_NETWORK -> bound_cRM x2 -> mask / subtraction -> iSTFT) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port), rank 0 only

A step = one pass of the hot path over one batch of synthetic utterances (BASELINE.json configs[1]:
batch 64 x 4 s, tensor-core mode, per GPU => weak scaling).  Rank 0 prints ONE JSON line.

Other BASELINE configurations (not the driver's default line; committed lines under profiles/bench_lines/):
    --variant dc | dr | drs          configs[2]: DC-Net (same net, S = Y (.) M) and the real-valued DR / DRS networks
    --workload train                 configs[4]: the training step (train-mode forward + losses + whole backward + NCCL gradient
                                     all-reduce + global-norm clip + Adam-amsgrad), batch 32 per GPU, hand-written kernels only
    --workload longform [--hours 1]  configs[3]: one hour of audio cut into independent 3.998 s windows, the window list
                                     sharded over the ranks (pipeline.shard_range), streamed host -> GPU -> host; a step = the
                                     whole hour, total work fixed => strong scaling
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "s_audio/s"
SR, HOP = 16000, 32
NET_NAME = {"dcs": "DCS-Net", "dc": "DC-Net", "dr": "DR-Net", "drs": "DRS-Net"}
HBM_SPEC_GBS = 8000.0     # the "about 8 TB/s" the north star names; fractions are reported against it AND the measured copy rate


def metric_name(variant):
    return f"seconds of 16 kHz audio enhanced/sec ({NET_NAME[variant]} fwd)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 = the tensor-core mode (tcgen05 kind::f16 on fp16 storage, fp32 accumulation; parity <= 2e-3)")
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--frames", type=int, default=2000, help="STFT frames per utterance (2000 = 3.998 s)")
    ap.add_argument("--variant", default="dcs", choices=["dcs", "dc", "dr", "drs"])
    ap.add_argument("--workload", default="batch", choices=["batch", "longform", "train"])
    ap.add_argument("--train-mode", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="train workload: tf32 = forward / dgrad convolutions on tcgen05 kind::tf32 (fp32 storage); bf16 = saved activations in bf16, forward "
                         "convolutions in kind::f16 on them (gradients / statistics / master weights fp32); fp32 = CUDA-core parity mode")
    ap.add_argument("--train-batch", type=int, default=32, help="train workload: utterances per GPU per step (BASELINE configs[4])")
    ap.add_argument("--hours", type=float, default=1.0, help="longform: hours of 16 kHz audio per step (whole job)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def audio_seconds(T):
    return HOP * (T - 1) / SR


# --------------------------------------------------------------------------------------------- FLOP model
def conv_flops_per_utterance(T, F=256, real=False):
    """Per layer {name: (dense, executed)} FLOPs of one utterance.  dense = the reference's formulation (2*M*N*K, 4 real
    MACs per complex MAC, every tap of the up-sampled input); executed = what the kernels' algorithm needs after the
    sub-pixel decomposition of nearest up-sampling + 3x3 conv (taps pre-summed per phase: /1.5 for (2,1), /2.25 for (2,2))."""
    ch = [1, 8, 16, 32, 64, 128, 128, 128]
    kE = [7, 7, 5, 5, 3, 3, 3]
    sE = [(2, 2), (2, 2), (2, 2), (2, 1), (2, 1), (2, 1), (2, 1)]
    up = [(2, 1), (2, 1), (2, 1), (2, 1), (2, 2), (2, 2), (2, 2)]
    H, W = F, T
    per_layer = {}
    for i in range(7):
        H, W = H // sE[i][0], W // sE[i][1]
        n, k = 2 * ch[i + 1], 2 * ch[i] * kE[i] ** 2
        if real and i == 0:
            k //= 2                                     # one real input channel (Appendix C: K = 49)
        f = 2.0 * (H * W) * n * k
        per_layer[f"enc{i}"] = (f, f)
    for i in range(7):
        cin = 2 * ch[7 - i]
        cout = ch[6 - i] if i < 6 else 1
        H, W = H * up[i][0], W * up[i][1]
        n = 2 * cout
        if real and i == 6:
            n //= 2                                     # one real output channel
        f = 2.0 * (H * W) * n * (2 * cin * 9)
        taps = (2 if up[i][0] == 2 else 3) * (2 if up[i][1] == 2 else 3)
        per_layer[f"dec{i}"] = (f, f * taps / 9.0)
    return per_layer


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15 and len(r) >= 8] or [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
            if any(r[col].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def is_real(variant):
    return variant in ("dr", "drs")


def make_weights(variant="dcs"):
    """Random-init weights of the reference architecture: C_NETWORK / R_NETWORK(config, hparams, seed=0).eval() (bit-identical
    to the reference's own constructors: tests/test_oracle_vs_reference.py, tests/test_rnet_oracle.py)."""
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network, r_network, config as cfg
    cls = r_network.R_NETWORK if is_real(variant) else c_network.C_NETWORK
    net = cls(cfg.config, cfg.hparams, 0).eval()
    return {k: v.detach() for k, v in net.state_dict().items()}


def oracle_enhance(variant):
    """The reference's CPU path (oracle port): fn(sd, noisy_audio) -> enhanced audio, STFT -> net -> combine -> iSTFT."""
    from oracle import dcsnet_oracle as O
    if is_real(variant):
        from oracle import rnet_oracle as RO
        return lambda sd, noisy: RO.enhance_spec(sd, O.stft(noisy), variant)["clean_audio"]
    return lambda sd, noisy: O.enhance_audio(sd, noisy, variant)["clean_audio"]


def cpu_reference_rate(sd, T, budget_s, variant, warmup=1, max_iters=50, min_iters=2):
    """The reference's CPU path (oracle port, all host threads) on a bounded sample: B=1 utterance of T frames per
    iteration, STFT -> net -> bound -> combine -> iSTFT.  Returns (audio s / s, iterations, seconds)."""
    from oracle import dcsnet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    fn = oracle_enhance(variant)
    _, _, noisy = O.synthetic_audio(1, HOP * (T - 1))
    with torch.no_grad():
        for _ in range(warmup):
            fn(sd, noisy)
        n, t0 = 0, time.perf_counter()
        while True:
            fn(sd, noisy)
            n += 1
            el = time.perf_counter() - t0
            if n >= max_iters or (n >= min_iters and el >= budget_s):
                break
    return n * audio_seconds(T) / el, n, el


def n_windows(args):
    return int(math.ceil(args.hours * 3600 * SR / (HOP * (args.frames - 1))))


def workload_text(args):
    T = args.frames
    if args.workload == "train":
        return (f"DCS-Net ({args.variant}) training step: train-mode forward (batch-statistic ComplexBatchNorm2d, dropout 0.1 / 0.2) + calc_loss + whole backward + gradient all-reduce + global-norm clip + Adam-amsgrad, batch {args.train_batch} x "
                f"{audio_seconds(T):.3f} s per GPU (T={T} frames), random-init weights seed 0 (BASELINE.json configs[4])")
    if args.workload == "longform":
        return (f"{NET_NAME[args.variant]} ({args.variant}) long-form inference: {args.hours:g} h of synthetic 16 kHz audio = "
                f"{n_windows(args)} independent windows of {audio_seconds(T):.3f} s (T={T} frames; last one zero-padded), window list "
                f"sharded contiguously over the GPUs, streamed pinned host -> GPU -> pinned host (BASELINE.json configs[3])")
    cfg_idx = "configs[1]" if args.variant == "dcs" else "configs[2]"
    return (f"{NET_NAME[args.variant]} ({args.variant}) batched inference, batch {args.batch} x {audio_seconds(T):.3f} s utterances per GPU "
            f"(T={T} frames, config.py STFT defaults), random-init weights seed 0 (BASELINE.json {cfg_idx})")


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path (the oracle port — /root/reference cannot
    travel to the GPU box and its dependencies are not installable), all host threads, each step a
    bounded sample (one utterance / window) of our arm's workload."""
    if rank != 0:
        return
    from oracle import dcsnet_oracle as O
    sd = make_weights(args.variant)
    torch.set_num_threads(os.cpu_count() or 1)
    fn = oracle_enhance(args.variant)
    _, _, noisy = O.synthetic_audio(1, HOP * (args.frames - 1))
    with torch.no_grad():
        for _ in range(args.warmup):
            fn(sd, noisy)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(sd, noisy)
        el = time.perf_counter() - t0
    val = args.steps * audio_seconds(args.frames) / el
    sample = f"1 utterance x {audio_seconds(args.frames):.3f} s per step (a bounded sample of the workload), fp32, torch CPU"
    line = {"impl": "reference", "metric": metric_name(args.variant), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "longform" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(args),
                       "reference_arm": "oracle port of the reference CPU path (oracle/dcsnet_oracle.py, oracle/rnet_oracle.py; "
                                        "/root/reference cannot travel)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def build_enhancer(args, sd, batch, n_samples):
    import dcsnet_b200 as D
    if is_real(args.variant):
        if not hasattr(D, "RealEnhancer"):
            raise SystemExit("the real-path (dr / drs) tensor-core plan is not built in this tree")
        return D.RealEnhancer(sd, batch=batch, n_samples=n_samples, mode=args.mode, variant=args.variant, graph=True)
    return D.Enhancer(sd, batch=batch, n_samples=n_samples, mode=args.mode, variant=args.variant, graph=True)


def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from dcsnet_b200 import pipeline
    from oracle import dcsnet_oracle as O  # synthetic audio generator + cpu_baseline leg only

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    T = args.frames
    n_samples = HOP * (T - 1)
    sd = make_weights(args.variant)
    longform = args.workload == "longform"
    nw = nw_total = 0
    if longform:
        # this rank's contiguous shard of the window list; the per-step batch is sized so that the shard divides evenly
        nw_total = n_windows(args)
        a_, b_ = pipeline.shard_range(nw_total, rank, world)
        nw = b_ - a_
        n_batches = max(1, math.ceil(nw / args.batch))
        B = max(2, 2 * math.ceil(nw / n_batches / 2))      # even: the LSTM's 4-sequence kernel needs 2 * B % 4 == 0
    else:
        B = args.batch
    enh = build_enhancer(args, sd, B, n_samples)
    plan = enh.plan
    if longform:
        g = torch.Generator().manual_seed(1234 + rank)
        wins = torch.empty(max(nw, 1), n_samples, dtype=torch.float32, pin_memory=True)
        wins.copy_(0.1 * torch.randn(max(nw, 1), n_samples, generator=g))
        outs = torch.empty(max(nw, 1), n_samples, dtype=torch.float32, pin_memory=True)
        wins_dev = wins.cuda()
        outs_dev = torch.empty_like(wins_dev)
    else:
        _, _, noisy = O.synthetic_audio(B, n_samples, seed=1234 + rank)
        enh.host_in.copy_(noisy)
        plan.audio_in.copy_(enh.host_in, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finalize=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        if finalize:
            finalize()   # e.g. make the timing stream wait for copies issued on side streams
        e1.record()
        barrier()
        w1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, w0, w1

    if longform:
        def step_dev():          # windows already resident in HBM, results stay in HBM
            for i in range(0, nw, B):
                n = min(B, nw - i)
                plan.audio_in[:n].copy_(wins_dev[i:i + n], non_blocking=True)
                plan.enhance_audio()
                outs_dev[i:i + n].copy_(plan.audio_out[:n], non_blocking=True)

        def step_e2e():          # pinned host windows -> GPU -> pinned host, copies overlapped with the neighbouring batches
            if nw:
                enh.enhance_windows_stream(wins[:nw], outs[:nw])
        step_serial = None
    else:
        step_dev, step_e2e, step_serial = plan.enhance_audio, enh.enhance_pinned_stream, enh.enhance_pinned

    # ---- device-resident throughput (inputs already in HBM)
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_dev, w0, w1 = timed(step_dev, args.steps)
    # ---- end to end through the public API: pinned host -> H2D -> graph -> D2H
    # (a) serial: H2D -> graph -> D2H on one stream; (b) streaming: the copies overlap the neighbouring steps' compute through
    # double-buffered staging (every step still copies its input from pinned host memory and its result back)
    ms_e2e_serial = None
    if step_serial is not None:
        for _ in range(args.warmup):
            step_serial()
        ms_e2e_serial, _, w1 = timed(step_serial, args.steps)
    for _ in range(args.warmup):
        step_e2e()
    enh.drain()
    ms_e2e, _, w1 = timed(step_e2e, args.steps, finalize=enh.drain)
    clocks = sampler.stop(w0, w1) if sampler else None

    if longform:
        audio_s_job = args.hours * 3600.0                       # the whole job's audio per step (all ranks together)
        h2d = d2h = nw * n_samples * 4
    else:
        audio_s_job = world * B * audio_seconds(T)
        h2d, d2h = enh.h2d_bytes, enh.d2h_bytes
    value = audio_s_job / (ms_dev / 1e3)
    e2e_value = audio_s_job / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel family (the tcgen05 implicit-GEMM conv), measured live with CUDA events on
    #      the launching stream over `steps` eager passes of the same plan (not under a profiler)
    roof, stage_ms = (None, None)
    if rank == 0:
        roof, stage_ms = measure_roofline(plan, args, B, T, clocks)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    launches_per_step = plan.graph_launches * (math.ceil(nw / B) if longform else 1)
    line = {
        "metric": metric_name(args.variant), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong" if longform else "weak", "vs_baseline": None,
        "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.mode], "data": "synthetic",
        "config": {"workload": workload_text(args),
                   "mode": args.mode, "global_batch": B * world, "samples_per_utterance": n_samples,
                   "parallelism": f"{'window' if longform else 'batch'}-sharded x{world}, no collective",
                   "l2": "working set per step (>= 3 GB of activations per 64 utterances) exceeds the 126 MB L2; no explicit flush",
                   "rtf": (ms_dev / 1e3) / audio_s_job},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "mode": ("Enhancer.enhance_windows_stream" if longform else "Enhancer.enhance_pinned_stream") +
                ": pinned host -> H2D -> graph -> D2H every batch, copies on their own streams overlapping the neighbouring batches "
                "(double-buffered staging)"},
        "gpu_launches": launches_per_step * args.steps,
        "kernels_per_step": launches_per_step,
        "clocks": clocks,
        "roofline": roof,
        "stage_ms": stage_ms,
    }
    if longform:
        line["config"].update({"windows_total": nw_total, "windows_this_rank": nw, "batch_per_step": B, "batches_per_step": math.ceil(nw / B)})
    if ms_e2e_serial is not None:
        line["e2e"].update({"ms_per_step_serial": ms_e2e_serial, "value_serial": audio_s_job / (ms_e2e_serial / 1e3)})
    if world == 1 and not args.no_cpu_baseline:
        v, n, el = cpu_reference_rate(sd, T, args.cpu_baseline_seconds, args.variant)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n} x (1 utterance of {audio_seconds(T):.3f} s) in {el:.1f} s: STFT->net->bound->combine->iSTFT, "
                                          "fp32 torch CPU, oracle port (oracle/)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_family_summary():
    """From the newest committed `ncu --set full` summary under profiles/ (tools/ncu_summary.py), for the tensor-core conv
    family (cconv_tc_kernel outside the LSTM projections + cconv_strip_kernel): mean dram__bytes_read.sum +
    dram__bytes_write.sum per launch, and the TIME-WEIGHTED sm__pipe_tensor_cycles_active %.  (None, None, None) when no
    summary is committed."""
    import csv
    import glob
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.csv")) if "_real_" not in os.path.basename(f) and "_train_" not in os.path.basename(f))
    if not files:
        return None, None, None
    rows = list(csv.reader(open(files[-1])))
    hdr = rows[0]
    try:
        ir = next(i for i, h in enumerate(hdr) if h.startswith("dram_read_MB"))
        iw = next(i for i, h in enumerate(hdr) if h.startswith("dram_write_MB"))
        it = next(i for i, h in enumerate(hdr) if h.startswith("time_us"))
        ip = next(i for i, h in enumerate(hdr) if h.startswith("tensor_pipe_pct"))
    except StopIteration:
        return None, None, None
    unit = lambda h: 1e6 if "Mbyte" in h else (1e9 if "Gbyte" in h else (1e3 if "Kbyte" in h else 1.0))   # noqa: E731
    in_lstm, vals, tw, tt = False, [], 0.0, 0.0
    for r in rows[1:]:
        k = r[1]
        if "lstm_deinterleave" in k:
            in_lstm = True
        elif "lstm_combine" in k:
            in_lstm = False
        elif ("cconv_tc_kernel" in k and not in_lstm) or "cconv_strip_kernel" in k:
            vals.append(float(r[ir]) * unit(hdr[ir]) + float(r[iw]) * unit(hdr[iw]))
            tw += float(r[it]) * float(r[ip])
            tt += float(r[it])
    if not vals:
        return None, None, None
    return sum(vals) / len(vals), (tw / tt if tt else None), os.path.relpath(files[-1], ROOT)


def measure_roofline(plan, args, B, T, clocks=None):
    """Eager (un-graphed) instrumented passes: CUDA events around every kernel family on the launching stream."""
    from dcsnet_b200 import ops
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    from dcsnet_b200 import _lib as L
    acc = {}
    orig = {}
    nlaunch = {}

    def wrap(name, key_fn):
        fn = getattr(ops, name)
        orig[name] = fn

        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = L.launch_count()
            e0.record()
            r = fn(*a, **k)
            e1.record()
            key = key_fn(*a, **k)
            acc.setdefault(key, []).append((e0, e1))
            nlaunch[key] = nlaunch.get(key, 0) + L.launch_count() - n0
            return r
        setattr(ops, name, inner)

    wrap("cconv", lambda pk, s0, s1, dst, use_tc=False, **k: "conv_tc" if use_tc else "conv_ffma")
    wrap("cconv_strip", lambda *a, **k: "conv_strip")
    for n in ("stft", "istft", "chan_pool", "chan_gate", "spat_stats", "spat_apply", "attention_fused", "attention_stream", "clstm",
              "mask_combine", "enc0", "dec6_tail", "real_attention", "real_attention_stream", "chan_max", "rlstm", "rlstm_tc", "mag_phase",
              "real_mask_combine"):
        if hasattr(ops, n):
            wrap(n, lambda *a, _n=n, **k: _n)
    steps = max(3, min(args.steps, 10))
    try:
        plan._enqueue_from_audio()
        torch.cuda.synchronize()
        acc.clear()
        nlaunch.clear()
        for _ in range(steps):
            plan._enqueue_from_audio()
        torch.cuda.synchronize()
    finally:
        for n, fn in orig.items():
            setattr(ops, n, fn)
    stage_ms = {k: sum(a.elapsed_time(b) for a, b in v) / steps for k, v in acc.items()}
    launches = {k: nlaunch.get(k, 0) // steps for k in acc}   # kernel launches (dcs_launch_count), not op calls
    fl = conv_flops_per_utterance(T, real=is_real(args.variant))
    hbm = peaks.get("hbm_gbs")
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"
    if not hbm:
        hbm, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    roof = None
    tc_mode = args.mode in ("fp16", "bf16")
    # ---- tensor-core convolution family (the only dense contractions of the path)
    tc_keys = [k for k in ("conv_tc", "conv_strip") if k in stage_ms]
    if tc_mode and tc_keys:
        # every conv layer runs on tcgen05 in this mode when enc0 / dec6 are on the strip kernel; otherwise they are
        # CUDA-core kernels ("enc0", "dec6_tail" stages) and their FLOPs are excluded
        layers = [k for k in fl if not ((k == "enc0" and "enc0" in stage_ms) or (k == "dec6" and "dec6_tail" in stage_ms))]
        flops = B * sum(fl[k][0] for k in layers)
        flops_exec = B * sum(fl[k][1] for k in layers)
        t_ms = sum(stage_ms[k] for k in tc_keys)
        n_tc = sum(launches[k] for k in tc_keys)
        burst, sustained = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
        # which measured peak applies: the burst figure when the timed region ran at (close to) the maximum SM clock — a
        # short step is not power-capped —, the sustained one when the clocks sat at the power-capped level
        at_max_clock = bool(clocks) and clocks.get("sm_mhz", 0) >= 0.97 * clocks.get("sm_max_mhz", 1e9)
        if burst and (at_max_clock or not sustained):
            peak, which = burst, "MEASURED_PEAKS.json bf16_tflops (burst: the timed region ran at the maximum SM clock, not power-capped)"
        elif sustained:
            peak, which = sustained, "MEASURED_PEAKS.json bf16_tflops_sustained (SM clock below maximum during the timed region)"
        else:
            peak, which = 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
        ach = flops / (t_ms / 1e3) / 1e12
        ach_exec = flops_exec / (t_ms / 1e3) / 1e12
        traffic, pipe_pct, ncu_src = ncu_family_summary()
        roof = {"bound": "tensor",
                "kernel": "tcgen05 implicit-GEMM convs: dcs::cconv_tc_kernel (enc3..enc6, dec0..dec3; fc + LSTM input projections) + "
                          "dcs::cconv_strip_kernel (enc0, enc1, enc2, dec4, dec5, dec6 + mask tail)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_unit": "bytes per launch (dram read + write, ncu)", "traffic_source": ncu_src, "peak_source": which,
                "algorithmic_flops_per_step": flops, "executed_flops_per_step": flops_exec,
                "achieved_executed": ach_exec, "frac_executed": ach_exec / peak,
                "frac_vs_burst": ach / burst if burst else None, "frac_vs_sustained": ach / sustained if sustained else None,
                "frac_executed_vs_burst": ach_exec / burst if burst else None,
                "tensor_pipe_pct": pipe_pct,
                "tensor_pipe_pct_source": (f"time-weighted sm__pipe_tensor_cycles_active over the family's launches in {ncu_src}" if ncu_src else None),
                "layers": layers, "avg_launch_ms": t_ms / n_tc, "launches_per_step": n_tc, "family_ms_per_step": t_ms,
                "note": "achieved / frac use the ALGORITHMIC FLOPs = the reference's dense formulation (SURVEY Appendix C: 4 real MACs per "
                        "complex MAC, every tap of the up-sampled input); *_executed use the FLOPs the kernels' algorithm performs after "
                        "the sub-pixel decomposition (decoder taps pre-summed: /1.5 and /2.25).  The fc and LSTM-projection launches "
                        "are inside the time but add no FLOPs to either numerator."}
    elif args.mode == "fp32":
        flops = B * sum(v[0] for v in fl.values())
        t = stage_ms.get("conv_ffma", 0.0) / 1e3
        ach = flops / t / 1e12 if t else 0.0
        roof = {"bound": "tensor", "kernel": "dcs::cconv_ffma_kernel (fp32 CUDA-core mode; no tensor-core roofline applies)",
                "achieved": ach, "peak": peaks.get("bf16_tflops_sustained", 1590.0), "unit": "TFLOP/s",
                "frac": ach / peaks.get("bf16_tflops_sustained", 1590.0), "traffic": None}
    # ---- bandwidth-bound stages: algorithmic bytes (each tensor moved once) / measured stage time
    esz = 2 if tc_mode else 4
    att = [t for t in plan.enc] + [t for t in plan.dec[:-1]]
    att_bytes = sum(t.numel() * t.element_size() for t in att)          # every attended tensor once
    stage_bytes = {
        "stft": B * (4 * 32 * (T - 1) + 8 * 256 * T) + (B * 256 * T * 2 * esz if "enc0" not in stage_ms else 0),
        "istft": B * (8 * 256 * T + 4 * 32 * (T - 1)),
        "spat_stats": att_bytes + sum(t.shape[0] * t.shape[1] * t.shape[2] * 16 for t in att),
        "spat_apply": 2 * att_bytes + sum(t.shape[0] * t.shape[1] * t.shape[2] * 16 for t in att),
        "attention_fused": 2 * att_bytes,
        "attention_stream": 2 * att_bytes,     # x once in, y once out
        "real_attention": 2 * att_bytes,       # the algorithmic minimum (a multi-pass kernel set reads x more than once)
        "real_attention_stream": 2 * att_bytes,
        "dec6_tail": B * (2 * 128 * (T // 2) * 16 * esz + 2 * 8 * 256 * T),
        "enc0": B * (8 * 256 * T + 128 * (T // 2) * 16 * esz),
    }
    stages = []
    for k, nbytes in stage_bytes.items():
        if k in stage_ms and stage_ms[k] > 0:
            gbs = nbytes / (stage_ms[k] / 1e3) / 1e9
            stages.append({"stage": k, "bound": "hbm", "ms": stage_ms[k], "launches": launches[k], "algorithmic_bytes": nbytes,
                           "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "frac_vs_8tbs": gbs / HBM_SPEC_GBS})
    for k in ("clstm", "rlstm", "rlstm_tc"):
        if k in stage_ms:
            stages.append({"stage": k, "bound": "latency (2 x S sequential recurrent steps; reported against neither roof)",
                           "ms": stage_ms[k], "launches": launches.get(k), "steps": 2 * 2 * (T // 8)})
    if roof is not None:
        roof["stages"] = stages
        roof["hbm_peak_source"] = hbm_src + f"; frac_vs_8tbs = against the {HBM_SPEC_GBS / 1e3:g} TB/s the north star names"
    return roof, stage_ms


def ncu_train_summary():
    """Newest committed `*_train_ncu_summary.csv` (tools/gpu_train_prof.sh): for the tcgen05 GEMM kernels of the training step
    (cconv_tc_kernel, wgrad_tc_kernel) the mean DRAM bytes per launch and the time-weighted tensor-pipe %."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_train_ncu_summary.csv")))
    if not files:
        return None, None, None
    rows = list(csv.reader(open(files[-1])))
    hdr = rows[0]
    try:
        ir = next(i for i, h in enumerate(hdr) if h.startswith("dram_read_MB"))
        iw = next(i for i, h in enumerate(hdr) if h.startswith("dram_write_MB"))
        it = next(i for i, h in enumerate(hdr) if h.startswith("time_us"))
        ip = next(i for i, h in enumerate(hdr) if h.startswith("tensor_pipe_pct"))
    except StopIteration:
        return None, None, None
    unit = lambda h: 1e6 if "Mbyte" in h else (1e9 if "Gbyte" in h else (1e3 if "Kbyte" in h else 1.0))   # noqa: E731
    vals, tw, tt = [], 0.0, 0.0
    for r in rows[1:]:
        if "cconv_tc_kernel" in r[1] or "wgrad_tc_kernel" in r[1]:
            vals.append(float(r[ir]) * unit(hdr[ir]) + float(r[iw]) * unit(hdr[iw]))
            tw += float(r[it]) * float(r[ip])
            tt += float(r[it])
    if not vals:
        return None, None, None
    return sum(vals) / len(vals), (tw / tt if tt else None), os.path.relpath(files[-1], ROOT)


def train_metric():
    return "seconds of 16 kHz audio per second through the DCS-Net training step (fwd + bwd + all-reduce + Adam-amsgrad)"


def run_reference_train(args, rank):
    """Reference arm of the training workload: the oracle's restatement of the reference's training step (train_batch_2_loss +
    autograd backward, oracle/train_oracle.py — pinned by tests/golden/train_step.pt) + torch.optim.Adam(amsgrad) on the host cores,
    one utterance per step (a bounded sample)."""
    if rank != 0:
        return
    from oracle import dcsnet_oracle as O, train_oracle as TO
    import dcsnet_b200  # noqa: F401
    from dcsnet_b200 import c_network, config as cfg
    torch.set_num_threads(os.cpu_count() or 1)
    hp = dict(cfg.hparams)
    net = c_network.C_NETWORK(cfg.config, hp, 0)
    params = {k for k, _ in net.named_parameters()}
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    clean, noise, noisy = O.synthetic_audio(1, HOP * (args.frames - 1))
    specs = [O.stft(t) for t in (noise, noisy, clean)]
    leaves = [sd[k].requires_grad_(True) for k in sorted(params)]
    opt = torch.optim.Adam(leaves, lr=hp["lr"], eps=hp["optim_eps"], weight_decay=hp["optim_weight_decay"], amsgrad=True)

    def step():
        r = TO.train_step(sd, *specs, params, "dcs", drop=TO.dropout_torch_stream(hp["dropout_conv"], hp["dropout_fc"]))
        for k, leaf in zip(sorted(params), leaves):
            leaf.grad = r["grads"].get(k)
        torch.nn.utils.clip_grad_norm_([l for l in leaves if l.grad is not None], hp["gradient_clip_val"])
        opt.step()
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = args.steps * audio_seconds(args.frames) / el
    sample = f"1 utterance x {audio_seconds(args.frames):.3f} s per step (a bounded sample), fp32, torch CPU autograd + Adam"
    print(json.dumps({"impl": "reference", "metric": train_metric(), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": workload_text(args), "reference_arm": "oracle restatement of the reference training step (oracle/train_oracle.py)"},
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_train(args, rank, local_rank, world):
    """BASELINE configs[4]: one training step per `step` on every rank (its own synthetic batch), gradients averaged over NCCL."""
    import torch.distributed as dist
    import dcsnet_b200 as D
    from dcsnet_b200 import c_network, config as cfg, ops, train_engine, train_ops as T
    from oracle import dcsnet_oracle as O  # synthetic audio generator only
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, Tn = args.train_batch, args.frames
    net = c_network.C_NETWORK(cfg.config, dict(cfg.hparams), 0).cuda().train()
    step = train_engine.TrainStep(net, args.variant, mode=args.train_mode, seed=rank).init_optimizer()
    clean, noise, noisy = O.synthetic_audio(B, HOP * (Tn - 1), seed=1234 + rank)
    dev_specs = [ops.stft(t.cuda()) for t in (noise, noisy, clean)]
    host_specs = [s.cpu().pin_memory() for s in dev_specs]
    stage = [torch.empty_like(s) for s in dev_specs]
    loss_host = torch.empty(1, dtype=torch.float32, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        w1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, w0, w1

    def step_dev():
        step.step(*dev_specs)

    # end to end: every step's batch (three spectrograms, what the reference's DataLoader yields) comes from pinned host memory and the
    # loss goes back to the host.  (a) serial: H2D, step, D2H on one stream; (b) streaming: the NEXT batch's H2D runs on a copy stream
    # into the other staging set while the current step computes (a pin_memory + non_blocking prefetcher), every step still copies its
    # own batch inside the timed region.
    def step_e2e_serial():
        for d, h in zip(stage, host_specs):
            d.copy_(h, non_blocking=True)
        out = step.step(*stage)
        loss_host.copy_(out["train_loss"].reshape(1), non_blocking=True)

    copy_stream = torch.cuda.Stream()
    sets = [stage, [torch.empty_like(s) for s in dev_specs]]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"k": 0}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for d, h in zip(sets[slot], host_specs):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = state["k"] % 2
        torch.cuda.current_stream().wait_event(ready[cur])
        out = step.step(*sets[cur])
        loss_host.copy_(out["train_loss"].reshape(1), non_blocking=True)
        consumed[cur].record()
        prefetch(1 - cur)
        state["k"] += 1

    for _ in range(args.warmup):
        step_dev()
    n0 = D._lib.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_dev, w0, w1 = timed(step_dev, args.steps)
    launches = (D._lib.launch_count() - n0) // args.steps
    for _ in range(args.warmup):
        step_e2e_serial()
    ms_e2e_serial, _, w1 = timed(step_e2e_serial, args.steps)
    consumed[0].record()
    consumed[1].record()
    prefetch(0)
    for _ in range(args.warmup):
        step_e2e()
    ms_e2e, _, w1 = timed(step_e2e, args.steps)
    torch.cuda.synchronize()
    clocks = sampler.stop(w0, w1) if sampler else None
    # ---- stage split + the convolution GEMM family (forward, dgrad, wgrad) timed with CUDA events around every call of one eager step
    fam = {"conv_fwd_dgrad": [], "conv_wgrad": []}
    roof = None

    def wrap(mod, name, key):
        orig = getattr(mod, name)

        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **k)
            e1.record()
            fam[key].append((e0, e1))
            return r
        setattr(mod, name, inner)
        return orig
    # EVERY rank runs the instrumented step (optimizer_step all-reduces: a rank-0-only call would dead-lock); rank 0 reports it
    o1, o2, o3 = wrap(ops, "cconv", "conv_fwd_dgrad"), wrap(T, "wgrad", "conv_wgrad"), wrap(T, "wgrad_tc16", "conv_wgrad")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    step.forward(*dev_specs)
    ev[1].record()
    step.backward()
    ev[2].record()
    step.optimizer_step()
    ev[3].record()
    barrier()
    ops.cconv, T.wgrad, T.wgrad_tc16 = o1, o2, o3
    if rank == 0:
        fam_ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in fam.items()}
        dense = sum(f for f, _ in conv_flops_per_utterance(Tn).values()) * B
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        bf16 = float(peaks.get("bf16_tflops", 1663.0))
        conv_ms = fam_ms["conv_fwd_dgrad"] + fam_ms["conv_wgrad"]
        ach = 3 * dense / (conv_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "convolution GEMM family of the step: forward + data gradient (dcs::cconv_tc_kernel, tcgen05 kind::tf32; few-channel "
                                             "layers on CUDA cores) and weight gradient (dcs::wgrad_tc_kernel, tcgen05 kind::f16 on bf16 operand copies; "
                                             "encoder[0] / decoder[6] on CUDA-core kernels)",
                "achieved": ach, "peak": bf16 / 2, "unit": "TFLOP/s", "frac": ach / (bf16 / 2),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 issues at half the kind::f16 rate)" if peaks else "fallback 1663 / 2",
                "flops_per_step": 3 * dense, "flop_basis": "dense formulation, forward + dgrad + wgrad = 3 x forward", "family_ms": fam_ms,
                "traffic": ncu_train_summary()[0], "tensor_pipe_pct": ncu_train_summary()[1], "ncu_summary": ncu_train_summary()[2],
                "stage_ms": {"forward": ev[0].elapsed_time(ev[1]), "backward": ev[1].elapsed_time(ev[2]), "allreduce_clip_adam": ev[2].elapsed_time(ev[3])}}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    audio = world * B * audio_seconds(Tn)
    h2d = sum(h.numel() * h.element_size() for h in host_specs)
    line = {"metric": train_metric(), "value": audio / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.train_mode], "data": "synthetic",
            "config": {"workload": workload_text(args), "mode": args.train_mode, "global_batch": B * world, "steps_per_s": 1e3 / ms_dev,
                       "parallelism": f"data-parallel x{world}: local BatchNorm statistics, flat fp32 gradient buckets all-reduced over NCCL, then the fused "
                                      "clip + Adam-amsgrad kernel on every rank",
                       "parameters": int(step.flat_param.numel()),
                       "l2": "activations saved per step (~8 GB at batch 32) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": audio / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step_serial": ms_e2e_serial, "value_serial": audio / (ms_e2e_serial / 1e3),
                    "mode": "TrainStep.step on a pinned-host batch of three spectrograms (what the reference's DataLoader yields): every step's H2D "
                            "inside the timed region, issued on a copy stream one step ahead into a second staging set (serial figures: one stream)"},
            "gpu_launches": launches * args.steps, "kernels_per_step": launches, "clocks": clocks, "roofline": roof}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank, local_rank, world = dist_env()
    if args.workload == "train":
        if args.impl == "reference":
            run_reference_train(args, rank)
        else:
            run_train(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
