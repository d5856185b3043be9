"""Per-stage parity report of the CUDA path against the oracle (developer tool; run on a B200 via gpurun).

    python tools/gpu_check.py stft|istft|fp32|bf16|tcunit|time [B] [T]

Each part is meant to run in its own process (a faulting kernel poisons the CUDA context).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dcsnet_b200 as D  # noqa: E402
from dcsnet_b200 import ops, packing  # noqa: E402
from oracle import dcsnet_oracle as O, synthetic_weights as SW  # noqa: E402


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.is_complex():
        a, b = torch.view_as_real(a), torch.view_as_real(b)
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def cl_to_nchw(t):
    """(B,H,W,C,2) real -> complex NCHW on CPU"""
    return torch.view_as_complex(t.detach().float().cpu().contiguous()).permute(0, 3, 1, 2)


def part_stft(B, T):
    L = 32 * (T - 1)
    _, _, noisy = O.synthetic_audio(B, L)
    ref = O.stft(noisy)
    got = ops.stft(noisy.cuda())
    torch.cuda.synchronize()
    print(json.dumps({"part": "stft", "B": B, "T": T, "rel": rel(got, ref)}))


def part_istft(B, T):
    g = torch.Generator().manual_seed(3)
    spec = torch.complex(torch.randn(B, 256, T, generator=g), torch.randn(B, 256, T, generator=g)) * 0.3
    ref = O.spec_to_wave(spec)
    for exact in (False, True):
        got = ops.istft(spec.cuda(), atan2_eps=O.HPARAMS["atan2_eps"], exact_polar=exact)
        torch.cuda.synchronize()
        print(json.dumps({"part": "istft", "exact": exact, "B": B, "T": T, "rel": rel(got, ref)}))


def part_net(mode, B, T):
    sd = SW.make_state_dict(0)
    L = 32 * (T - 1)
    _, _, noisy = O.synthetic_audio(B, L)
    taps = {}
    spec = O.stft(noisy)
    ref = O.enhance_spec(sd, spec, taps=taps)
    ref_audio = O.spec_to_wave(ref["clean_spec"])
    pk = D.PackedNet(sd, "cuda", mode)
    plan = D.ForwardPlan(pk, B, T, keep_taps=True)
    out = plan.enhance_audio(noisy.cuda())
    torch.cuda.synchronize()
    res = {"part": mode, "B": B, "T": T}
    res["Y"] = rel(plan.Y, spec)
    res["bn0"] = rel(cl_to_nchw(plan.bn0), taps["bn0"])
    for k, v in plan.taps.items():
        if k in ("lstm", "fc"):
            t = torch.view_as_complex(v.detach().float().cpu().contiguous())
            t = t.reshape(B, -1, 128)
            res[k] = rel(t, taps[k])
        else:
            res[k] = rel(cl_to_nchw(v), taps[k])
    squeeze = lambda t: t.squeeze(0) if B == 1 else t
    res["net_out"] = rel(squeeze(plan.net_out), ref["net_out"])
    res["mask"] = rel(squeeze(plan.mask), ref["mask"])
    res["clean_spec"] = rel(plan.clean_spec, ref["clean_spec"])
    res["noise_spec"] = rel(plan.noise_spec, ref["noise_spec"])
    res["clean_audio"] = rel(out, ref_audio)
    clean = O.synthetic_audio(B, L)[0]
    res["d_sisdr_db"] = abs(float(O.si_snr(clean, out.cpu()) - O.si_snr(clean, ref_audio)))
    print(json.dumps(res))


def part_tcunit(B, T):
    """tcgen05 conv vs the FFMA conv on identical bf16 inputs, layer by layer (isolates the tensor-core kernel)."""
    sd = SW.make_state_dict(0)
    pk = D.PackedNet(sd, "cuda", "bf16")
    g = torch.Generator().manual_seed(5)
    H, W = 256, T
    shapes = []
    for i in range(7):
        cin = pk.enc[i].cin
        shapes.append(("enc%d" % i, pk.enc[i], (B, H, W, cin), None))
        H, W = ops.conv_out_hw(pk.enc[i], H, W)
    Hs, Ws = [256, 128, 64, 32, 16, 8, 4, 2], None
    H, W = 2, T // 8
    for i in range(7):
        c = pk.dec[i].cin // 2
        shapes.append(("dec%d" % i, pk.dec[i], (B, H, W, c), (B, H, W, c)))
        H, W = H * pk.dec[i].up[0], W * pk.dec[i].up[1]
    for name, p, s0, s1 in shapes:
        if (2 * p.cin) % 16 or (2 * s0[3]) % 16:
            print(json.dumps({"part": "tcunit", "layer": name, "skipped": "2*Cin not a multiple of 16"}))
            continue
        x0 = torch.randn(*s0, 2, generator=g).cuda().bfloat16()
        x1 = torch.randn(*s1, 2, generator=g).cuda().bfloat16() if s1 else None
        oh, ow = ops.conv_out_hw(p, s0[1], s0[2])
        ref = torch.empty(B, oh, ow, p.cout, 2, device="cuda", dtype=torch.float32)
        got = torch.full((B, oh, ow, p.cout, 2), float("nan"), device="cuda", dtype=torch.float32 if name == "dec6" else torch.bfloat16)
        ops.cconv(p, x0, x1, ref, use_tc=False)
        # the FFMA reference uses fp32 weights; round them like the tensor-core operand for an apples-to-apples check
        ops.cconv(p, x0, x1, got, use_tc=True)
        torch.cuda.synchronize()
        print(json.dumps({"part": "tcunit", "layer": name, "rel": rel(got.float(), ref), "nan": int(torch.isnan(got.float()).sum())}))


def part_stripunit(B, T):
    """row-strip tcgen05 conv vs the FFMA conv on identical bf16 inputs + timing against the general tcgen05 kernel"""
    sd = SW.make_state_dict(0)
    pk = D.PackedNet(sd, "cuda", "bf16")
    g = torch.Generator().manual_seed(5)
    cases = [("enc1", pk.enc[1], (B, 128, T // 2, 8), None, True, 1), ("dec4", pk.dec[4], (B, 32, T // 8, 32), True, False, 2),
             ("dec4m", pk.dec[4], (B, 32, T // 8, 32), True, True, 2), ("dec5", pk.dec[5], (B, 64, T // 4, 16), True, True, 1)]
    # encoder[0]: Toeplitz strip kernel vs the FFMA enc0 kernel (which reads the fp32 spectrogram)
    x = torch.randn(B, 256, T, 1, 2, generator=g).cuda().bfloat16()
    sp0 = packing.StripEnc0(pk.enc[0], device="cuda")
    ref0 = torch.empty(B, 128, T // 2, 8, 2, device="cuda")
    got0 = torch.empty(B, 128, T // 2, 8, 2, device="cuda", dtype=torch.bfloat16)
    ops.cconv(pk.enc[0], x, None, ref0, use_tc=False)
    ops.cconv_strip(sp0, packing.StripEnc0.view_src(x), None, got0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.cconv_strip(sp0, packing.StripEnc0.view_src(x), None, got0)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"part": "stripunit", "layer": "enc0", "rel": rel(got0.float(), ref0), "ms_strip": round(e0.elapsed_time(e1) / 5, 4)}), flush=True)
    del x, ref0, got0
    for name, p, s0, two, merged, groups in cases:
        x0 = torch.randn(*s0, 2, generator=g).cuda().bfloat16()
        x1 = torch.randn(*s0, 2, generator=g).cuda().bfloat16() if two else None
        c0 = s0[3]
        sp = packing.StripConv(p, c0, c0 if two else 0, merged=merged, groups=groups, device="cuda")
        oh, ow = ops.conv_out_hw(p, s0[1], s0[2])
        ref = torch.empty(B, oh, ow, p.cout, 2, device="cuda", dtype=torch.bfloat16)
        got = torch.full((B, oh, ow, p.cout, 2), float("nan"), device="cuda", dtype=torch.bfloat16)
        pool_ref = torch.zeros(B, p.cout, 2, device="cuda")
        pool = torch.zeros(B, p.cout, 2, device="cuda")
        ops.cconv(p, x0, x1, ref, use_tc=True, pool_sums=pool_ref)
        ops.cconv_strip(sp, x0, x1, got, pool_sums=pool)
        torch.cuda.synchronize()
        res = {"part": "stripunit", "layer": name, "rel": rel(got.float(), ref.float()), "pool_rel": rel(pool, pool_ref),
               "nan": int(torch.isnan(got.float()).sum())}
        for label, fn in (("ms_strip", lambda: ops.cconv_strip(sp, x0, x1, got)), ("ms_tc", lambda: ops.cconv(p, x0, x1, ref, use_tc=True))):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[label] = round(e0.elapsed_time(e1) / 5, 4)
        print(json.dumps(res), flush=True)


def part_attn(B, T):
    """Streaming attention kernel vs the separate stats/apply kernels on every attended-tensor geometry: parity + time."""
    sd = SW.make_state_dict(1)
    F = 256
    shapes = [(8, F // 2, T // 2), (16, F // 4, T // 4), (32, F // 8, T // 8), (64, F // 16, T // 8), (128, F // 32, T // 8),
              (128, F // 64, T // 8), (128, F // 128, T // 8)]
    for C, H, W in shapes:
        i = {128: 0, 64: 3, 32: 4, 16: 5, 8: 6}[C]
        ca = packing.pack_channel_attention(sd, f"skip_attention.{2 * i}.", "cuda")
        w7 = packing.pack_spatial_attention(sd, f"skip_attention.{2 * i + 1}.", "cuda")
        g = torch.Generator().manual_seed(C + H)
        x = torch.randn(B, H, W, C, 2, generator=g).to(torch.bfloat16).cuda()
        sums = torch.zeros(B, C, 2, device="cuda")
        ops.chan_pool(x, sums)
        gate = torch.empty(B, C, 2, device="cuda")
        stats = torch.empty(B, H * W, 4, device="cuda")
        ref = torch.empty(B, H, W, C, 2, device="cuda")
        ref16 = torch.empty_like(x)
        got = torch.full_like(x, float("nan"))

        def sep(out):
            ops.spat_stats(x, None, stats, sums=sums, ca=ca, gate_out=gate)
            ops.spat_apply(x, gate, stats, w7, out)
        sep(ref)
        ops.attention_stream(x, sums, ca, w7, got)
        torch.cuda.synchronize()
        nan = int(torch.isnan(got.float()).sum())
        r = float((got.float() - ref).abs().max() / ref.abs().max())

        def timeit(fn, n=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        ms_s = timeit(lambda: ops.attention_stream(x, sums, ca, w7, got))
        ms_r = timeit(lambda: sep(ref16))
        mb = x.numel() * 2 / 1e6
        print(json.dumps({"part": "attn", "C": C, "H": H, "W": W, "rel": r, "nan": nan, "ms_stream": round(ms_s, 4),
                          "ms_separate": round(ms_r, 4), "GBps_stream": round(2 * mb / ms_s, 1)}), flush=True)


def part_layers(B, T):
    """per-layer timing of the tcgen05 conv (and the FFMA layers) at full size"""
    sd = SW.make_state_dict(0)
    pk = D.PackedNet(sd, "cuda", "fp16")
    plan = D.ForwardPlan(pk, B, T, want_aux=False)
    _, _, noisy = O.synthetic_audio(B, 32 * (T - 1))
    plan.enhance_audio(noisy.cuda())
    torch.cuda.synchronize()
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    fl = bench.conv_flops_per_utterance(T)
    Lr = 7
    jobs = []
    x = plan.bn0
    for i in range(Lr):
        jobs.append((f"enc{i}", pk.enc[i], x, None, plan.enc[i]))
        x = plan.enc[i]
    d = plan.fc
    for i in range(Lr):
        jobs.append((f"dec{i}", pk.dec[i], d, plan.skip[i], plan.dec[i]))
        d = plan.datt[i]
    def timeit(label, fn, extra=None):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        d = {"op": label, "ms": round(e0.elapsed_time(e1) / 5, 4)}
        d.update(extra or {})
        print(json.dumps(d))

    Hh, Ww = plan.enc[6].shape[1], plan.enc[6].shape[2]
    timeit("stft", lambda: ops.stft(plan.audio_in, plan.Y))
    timeit("enc0", lambda: ops.enc0(pk.enc[0], plan.Y, pk.bn0, plan.enc[0]))
    timeit("clstm", lambda: ops.clstm(plan.enc[6].view(B, Hh * Ww, 128, 2), plan.lat.view(B, Hh * Ww, 128, 2), pk.lstm, plan.lstm_ws, use_tc=True))
    timeit("fc", lambda: ops.cconv(pk.fc, plan.lat.view(B, 1, Hh * Ww, 128, 2), None, plan.fc.view(B, 1, Hh * Ww, 128, 2), use_tc=True))
    timeit("dec6_tail", lambda: ops.dec6_tail(pk.dec[6], plan.datt[5], plan.skip[6], plan.Y, plan.clean_spec))
    timeit("istft", lambda: ops.istft(plan.clean_spec, plan.audio_out))
    for i in range(7):
        x = plan.enc[6 - i]
        timeit(f"skip_att{i}", lambda: plan._attention(x, pk.skip_ca[i], pk.skip_sa[i], plan.skip[i]),
               {"C": x.shape[3], "MB": round(x.numel() * x.element_size() / 1e6, 1)})
    for i in range(6):
        x = plan.dec[i]
        timeit(f"dec_att{i}", lambda: plan._attention(x, pk.dec_ca[i], pk.dec_sa[i], plan.datt[i]),
               {"C": x.shape[3], "MB": round(x.numel() * x.element_size() / 1e6, 1)})
    x = plan.dec[5]
    st = plan.stats.view(-1)[: B * x.shape[1] * x.shape[2] * 4].view(B, -1, 4)
    gt = plan.gate.view(-1)[: B * 8 * 2].view(B, 8, 2)
    timeit("dec_att5.stats", lambda: ops.spat_stats(x, gt, st))
    timeit("dec_att5.apply", lambda: ops.spat_apply(x, gt, st, pk.dec_sa[5], plan.datt[5]))
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    for name, p, s0, s1, dst in jobs:
        if name in ("enc0", "dec6"):
            continue
        use_tc = (2 * p.cin) % 16 == 0
        D._lib.lib().dcs_tc_set_debug_buffer(D._lib.ptr(dbg))
        dbg.zero_()
        ops.cconv(p, s0, s1, dst, use_tc=use_tc)
        torch.cuda.synchronize()
        D._lib.lib().dcs_tc_set_debug_buffer(None)
        d = dbg.view(148, 8).double()
        act = d[:, 7] > 0
        m = d[act].mean(0)
        print(json.dumps({"layer": name, "dbg_kcycles": {"prod_wait_empty": round(float(m[0]) / 1e3, 1), "prod_total": round(float(m[1]) / 1e3, 1),
              "mma_wait_full": round(float(m[2]) / 1e3, 1), "mma_wait_acc": round(float(m[3]) / 1e3, 1), "mma_total": round(float(m[4]) / 1e3, 1),
              "epi_wait": round(float(m[5]) / 1e3, 1), "epi_total": round(float(m[6]) / 1e3, 1), "tiles_per_cta": round(float(m[7]), 1)}}))
        for _ in range(2):
            ops.cconv(p, s0, s1, dst, use_tc=use_tc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            ops.cconv(p, s0, s1, dst, use_tc=use_tc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        in_bytes = s0.numel() * s0.element_size() + (s1.numel() * s1.element_size() if s1 is not None else 0)
        out_bytes = dst.numel() * dst.element_size()
        print(json.dumps({"layer": name, "tc": use_tc, "ms": round(ms, 4), "tflops_ref": round(B * fl[name][0] / ms / 1e9, 1), "tflops_exec": round(B * fl[name][1] / ms / 1e9, 1),
                          "N": 2 * p.cout, "K": p.ntaps * 2 * p.cin * p.phases, "in_MB": round(in_bytes / 1e6, 1), "out_MB": round(out_bytes / 1e6, 1),
                          "GBps_min": round((in_bytes + out_bytes) / ms / 1e6, 1)}))


def part_time(mode, B, T):
    sd = SW.make_state_dict(0)
    pk = D.PackedNet(sd, "cuda", mode)
    plan = D.ForwardPlan(pk, B, T, want_aux=False)
    _, _, noisy = O.synthetic_audio(B, 32 * (T - 1))
    plan.audio_in.copy_(noisy.cuda())
    for _ in range(2):
        plan.enhance_audio()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        plan.enhance_audio()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"part": "time", "mode": mode, "B": B, "T": T, "ms": ms, "audio_s_per_s": B * O.audio_seconds(T) / (ms / 1e3)}))
    plan.capture()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        plan.enhance_audio()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"part": "time_graph", "mode": mode, "B": B, "T": T, "ms": ms, "launches": plan.graph_launches,
                      "audio_s_per_s": B * O.audio_seconds(T) / (ms / 1e3)}))


if __name__ == "__main__":
    part = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    t0 = time.time()
    if part == "stft":
        part_stft(B, T)
    elif part == "istft":
        part_istft(B, T)
    elif part in ("fp32", "bf16"):
        part_net(part, B, T)
    elif part == "tcunit":
        part_tcunit(B, T)
    elif part == "attn":
        part_attn(B, T)
    elif part == "stripunit":
        part_stripunit(B, T)
    elif part == "layers":
        part_layers(B, T)
    elif part.startswith("time"):
        part_time(part.split("_")[1], B, T)
    print(f"# {part} done in {time.time() - t0:.1f}s", file=sys.stderr)
