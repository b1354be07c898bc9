"""Micro-benchmarks of the individual kernels (CUDA events, L2 flushed between iterations).
Usage: python tools/bench_kernels.py [gemm] [compress] [splice]   -> JSON lines on stdout."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402

PEAKS = {"hbm_gbs": 6543.7, "bf16_tflops": 1675.7}
try:
    PEAKS.update(json.load(open(__file__.rsplit("/tools/", 1)[0] + "/MEASURED_PEAKS.json")))
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=20, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_gemm():
    shapes = [(4096, 3072, 2048), (4096, 2048, 2048), (4096, 16384, 2048), (4096, 2048, 8192), (8192, 8192, 8192),
              (15616, 3072, 2048), (2304, 128256, 2048), (64, 3072, 2048)]
    for M, N, K in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        for bn in (128, 256):
            med, best = timeit(lambda: ops.gemm(a, b, block_n=bn))
            tf = 2.0 * M * N * K / (med * 1e-3) / 1e12
            print(json.dumps({"kernel": "gemm_tcgen05", "M": M, "N": N, "K": K, "block_n": bn, "ms": round(med, 4),
                              "ms_best": round(best, 4), "tflops": round(tf, 1),
                              "frac_of_measured_peak": round(tf / PEAKS["bf16_tflops"], 3)}), flush=True)
        med, best = timeit(lambda: torch.matmul(a, b.t()))
        tf = 2.0 * M * N * K / (med * 1e-3) / 1e12
        print(json.dumps({"kernel": "torch.matmul(cuBLAS)", "M": M, "N": N, "K": K, "ms": round(med, 4),
                          "tflops": round(tf, 1)}), flush=True)


def bench_compress():
    for B, T, n_tok, D, rate in [(64, 1500, 800, 1024, 4), (64, 1500, 800, 1024, 16), (64, 400, 400, 1024, 2),
                                 (64, 400, 400, 1024, 5)]:
        x = torch.randn(B, T, D, device="cuda").bfloat16()
        for mode in ("avg-pooling", "stack"):
            med, best = timeit(lambda: ops.matryoshka_compress(x, n_tok, rate, mode))
            n_out = n_tok // rate
            byts = B * n_out * rate * D * 2 + B * n_out * (D if mode != "stack" else D * rate) * 2
            gbs = byts / (med * 1e-3) / 1e9
            print(json.dumps({"kernel": "matryoshka_compress", "mode": mode, "B": B, "n_tok": n_tok, "rate": rate,
                              "ms": round(med, 4), "GBs": round(gbs, 1),
                              "frac_of_measured_hbm": round(gbs / PEAKS["hbm_gbs"], 3)}), flush=True)


def bench_splice():
    H, V, L = 2048, 128261, 48
    for B, n_a, n_v in [(64, 200, 200), (64, 50, 80), (16, 200, 200)]:
        embed = torch.randn(V, H, device="cuda").bfloat16()
        tokens = torch.randint(0, V, (B, L), device="cuda")
        a = torch.randn(B, n_a, H, device="cuda").bfloat16()
        v = torch.randn(B, n_v, H, device="cuda").bfloat16()
        prompts = [torch.randn(p, H, device="cuda").bfloat16() for p in (6, 6, 8)]
        lay = ops.SpliceLayout(tokens=tokens, labels=tokens, embed=embed, audio_tok=a, video_tok=v, prompts=prompts,
                               marker_ids=(V - 4, V - 3, V - 2, V - 1), has_bos=True)
        outs = [torch.empty(B, s, H, device="cuda", dtype=torch.bfloat16) for s in lay.seq_len]
        outl = [torch.empty(B, s, device="cuda", dtype=torch.int64) for s in lay.seq_len]
        med, best = timeit(lambda: ops.splice_prompt(lay, outs, outl))
        rows = B * sum(lay.seq_len)
        # algorithmic bytes: every destination row written once (H*2) + its source row read once, media rows
        # are read once but written twice (own task + AVSR)  => reads = unique source rows
        uniq = B * (n_a + n_v) + B * (L + 4 + 4) + 20
        byts = rows * H * 2 + rows * 8 + (B * (n_a + n_v) + rows - 2 * B * (n_a + n_v)) * H * 2
        gbs = byts / (med * 1e-3) / 1e9
        print(json.dumps({"kernel": "splice_prompt", "B": B, "n_a": n_a, "n_v": n_v, "rows": rows,
                          "ms": round(med, 4), "GBs": round(gbs, 1),
                          "frac_of_measured_hbm": round(gbs / PEAKS["hbm_gbs"], 3)}), flush=True)


def bench_attention():
    """Path shapes at B=16: Whisper-medium (S=1500, non-causal), AV-HuBERT (S=400), Llama-1B segments (causal GQA),
    plus the head_dim-128 shapes of Qwen2.5-3B / Llama-3.1-8B; cuDNN SDPA timed beside each."""
    import torch.nn.functional as F
    shapes = [("whisper-m", 16, 1500, 16, 16, 64, False), ("avhubert-l", 16, 400, 16, 16, 64, False),
              ("llama1b-avsr", 16, 460, 32, 8, 64, True), ("llama1b-asr", 16, 256, 32, 8, 64, True),
              ("qwen3b-avsr", 16, 460, 16, 2, 128, True), ("llama8b-avsr", 16, 460, 32, 8, 128, True)]
    for name, B, S, nh, nkv, hd, causal in shapes:
        M = B * S
        qkv = torch.randn(M, (nh + 2 * nkv) * hd, device="cuda").bfloat16()
        out = torch.empty(M, nh * hd, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(nh, M, device="cuda")
        dout = torch.randn(M, nh * hd, device="cuda").bfloat16()
        dqkv = torch.empty_like(qkv)
        seg = [(0, B, S, 0)]
        fl = 4.0 * B * nh * S * S * hd * (0.5 if causal else 1.0)
        med, _ = timeit(lambda: ops.attention_fwd(qkv, out, seg, nh, nkv, hd, causal, lse=lse))
        medb, _ = timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, dqkv, seg, nh, nkv, hd, causal))
        blk = qkv.view(B, S, -1)
        q = blk[..., : nh * hd].view(B, S, nh, hd).transpose(1, 2).detach().requires_grad_(True)
        k = blk[..., nh * hd: (nh + nkv) * hd].view(B, S, nkv, hd).transpose(1, 2).detach().requires_grad_(True)
        v = blk[..., (nh + nkv) * hd:].view(B, S, nkv, hd).transpose(1, 2).detach().requires_grad_(True)
        medl, _ = timeit(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=causal, enable_gqa=nh != nkv))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=causal, enable_gqa=nh != nkv)
        do = dout.view(B, S, nh, hd).transpose(1, 2)
        medlb, _ = timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
        print(json.dumps({"kernel": "attention", "shape": name, "B": B, "S": S, "heads": nh, "kv_heads": nkv, "hd": hd,
                          "causal": causal, "fwd_ms": round(med, 4), "fwd_tflops": round(fl / med / 1e9, 1),
                          "bwd_ms": round(medb, 4), "bwd_tflops": round(2.5 * fl / medb / 1e9, 1),
                          "sdpa_fwd_ms": round(medl, 4), "sdpa_bwd_ms": round(medlb, 4)}), flush=True)


def bench_transforms():
    """Input-pipeline kernels on one 16 s utterance (400 RGB 96x96 frames, 256000 samples), train pipelines."""
    from omni_avsr_b200 import transforms as ptr
    video = torch.randint(0, 256, (400, 3, 96, 96), dtype=torch.uint8, device="cuda")
    vt = ptr.VideoTransform("train", out_dtype=torch.bfloat16)
    med, _ = timeit(lambda: vt(video))
    byts = 400 * 3 * 88 * 88 + 400 * 88 * 88 * 2
    print(json.dumps({"kernel": "video_transform(train, bf16 out)", "frames": 400, "ms": round(med, 4),
                      "GBs": round(byts / med / 1e6, 1), "note": "includes the host-side RNG draws of the mirror"}), flush=True)
    wave = torch.randn(256000, 1, device="cuda") * 0.1
    at = ptr.AudioTransform("train", noise=torch.randn(1, 400000) * 0.3)
    med, _ = timeit(lambda: at(wave))
    byts = 256000 * 4 * (2 + 2 + 1)
    print(json.dumps({"kernel": "audio_transform(train)", "samples": 256000, "ms": round(med, 4),
                      "GBs": round(byts / med / 1e6, 1), "note": "two passes (sums, apply); includes host-side RNG draws"}),
          flush=True)


def fused_setup(B, ra=4, rv=2, H=2048, I=2048, D=1024, L=48, V=128261, infer_task=None):
    """Inputs of the compression -> projector -> splice stage at BASELINE config 2 geometry (16 s clips)."""
    g = torch.Generator(device="cuda").manual_seed(0)
    xa = torch.randn(B, 1500, D, device="cuda", generator=g).bfloat16()
    xv = torch.randn(B, 400, D, device="cuda", generator=g).bfloat16()

    def proj(K1):
        return [(torch.randn(I, K1, device="cuda", generator=g) / K1 ** 0.5).bfloat16(),
                (torch.randn(I, device="cuda", generator=g) * 0.1).bfloat16(),
                (torch.randn(H, I, device="cuda", generator=g) / I ** 0.5).bfloat16(),
                (torch.randn(H, device="cuda", generator=g) * 0.1).bfloat16()]
    pa, pv = proj(D), proj(D)
    embed = torch.randn(V, H, device="cuda", generator=g).bfloat16()
    tokens = torch.randint(0, V - 8, (B, L), device="cuda", generator=g)
    prompts = [torch.randn(pl, H, device="cuda", generator=g).bfloat16() for pl in (6, 6, 8)]
    marker = (V - 4, V - 3, V - 2, V - 1)
    return dict(xa=xa, xv=xv, pa=pa, pv=pv, embed=embed, tokens=tokens, prompts=prompts, marker=marker, ra=ra, rv=rv,
                na=800 // ra, nv=400 // rv, B=B, H=H, I=I, D=D, L=L)


def fused_bytes_flops(c):
    """Algorithmic bytes / flops of the stage: encoder rows read once, each projected token written twice (own
    task + AVSR), marker / prompt / text rows read + written, labels, projector weights once; 2*n*(D*I + I*H) flops."""
    B, H, I, D, L = c["B"], c["H"], c["I"], c["D"], c["L"]
    na, nv, ra, rv = c["na"], c["nv"], c["ra"], c["rv"]
    media = B * (na * ra + nv * rv) * D * 2 + 2 * B * (na + nv) * H * 2
    S = [1 + na + 2 + 6 + L - 1, 1 + nv + 2 + 6 + L - 1, 1 + na + 2 + nv + 2 + 8 + L - 1]
    other_rows = B * (sum(S) - 2 * (na + nv))
    other = other_rows * H * 2 * 2 + B * sum(S) * 8
    weights = 2 * (D * I + I + I * H + H) * 2
    flops = 2.0 * B * (na + nv) * (D * I + I * H)
    return media + other, weights, flops


def bench_fused():
    for B in (32, 64):
        for ra, rv in ((4, 2), (16, 5)):
            c = fused_setup(B, ra, rv)
            lay_f = ops.SpliceLayout(tokens=c["tokens"], labels=c["tokens"], embed=c["embed"], audio_tok=None, video_tok=None,
                                     prompts=c["prompts"], marker_ids=c["marker"], has_bos=True, n_audio=c["na"],
                                     n_video=c["nv"])
            outs = [torch.empty(B, s, c["H"], device="cuda", dtype=torch.bfloat16) for s in lay_f.seq_len]
            outl = [torch.empty(B, s, device="cuda", dtype=torch.int64) for s in lay_f.seq_len]
            a_in = ops.PoolProjectInput(c["xa"], 800, ra, *c["pa"])
            v_in = ops.PoolProjectInput(c["xv"], 400, rv, *c["pv"])
            med, best = timeit(lambda: ops.pool_project_splice(lay_f, outs, outl, a_in, v_in, "avg-pooling"))

            def unfused():
                ca = ops.matryoshka_compress(c["xa"], 800, ra, "avg-pooling")
                cv = ops.matryoshka_compress(c["xv"], 400, rv, "avg-pooling")
                ta = ops.gemm(ops.gemm(ca.view(-1, c["D"]), c["pa"][0], bias=c["pa"][1], act="relu"), c["pa"][2], bias=c["pa"][3])
                tv = ops.gemm(ops.gemm(cv.view(-1, c["D"]), c["pv"][0], bias=c["pv"][1], act="relu"), c["pv"][2], bias=c["pv"][3])
                lay = ops.SpliceLayout(tokens=c["tokens"], labels=c["tokens"], embed=c["embed"],
                                       audio_tok=ta.view(B, c["na"], -1), video_tok=tv.view(B, c["nv"], -1),
                                       prompts=c["prompts"], marker_ids=c["marker"], has_bos=True)
                ops.splice_prompt(lay, outs, outl)
            medu, bestu = timeit(unfused)
            byts, wbytes, flops = fused_bytes_flops(c)
            t_hbm = (byts + wbytes) / (PEAKS["hbm_gbs"] * 1e9)
            t_tc = flops / (PEAKS["bf16_tflops"] * 1e12)
            print(json.dumps({"kernel": "pool_project_splice (fused, 1 launch)", "B": B, "rates": [ra, rv],
                              "ms": round(med, 4), "ms_best": round(best, 4), "unfused_7_launches_ms": round(medu, 4),
                              "MB_per_utt": round(byts / B / 1e6, 3), "GFLOP_per_utt": round(flops / B / 1e9, 3),
                              "tflops": round(flops / (med * 1e-3) / 1e12, 1),
                              "frac_tensor_burst": round(flops / (med * 1e-3) / 1e12 / PEAKS["bf16_tflops"], 3),
                              "GBs": round((byts + wbytes) / med / 1e6, 1),
                              "bound_ms": round(1e3 * max(t_hbm, t_tc), 4),
                              "frac_of_bound": round(1e3 * max(t_hbm, t_tc) / med, 3)}), flush=True)


def bench_skinny():
    """Decode-step GEMMs (64 token rows): weight-streaming kernel vs the general kernel, GB/s of weight bytes."""
    import os
    for M, N, K in ((64, 2048, 2048), (64, 2048, 8192), (64, 3072, 2048), (64, 16384, 2048), (64, 256, 2048), (15, 2048, 8192)):
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        res = torch.randn(M, N, device="cuda").bfloat16()
        med, best = timeit(lambda: ops.gemm(x, w, residual=res, skinny=True))
        medo, _ = timeit(lambda: ops.gemm(x, w, residual=res, block_n=256))
        print(json.dumps({"kernel": "gemm_skinny", "M": M, "N": N, "K": K, "split_env": os.environ.get("OMNI_SKINNY_SPLIT"),
                          "us": round(med * 1e3, 2), "us_best": round(best * 1e3, 2), "GBs": round(N * K * 2 / med / 1e6, 1),
                          "frac_hbm": round(N * K * 2 / med / 1e6 / PEAKS["hbm_gbs"], 3),
                          "general_kernel_us": round(medo * 1e3, 2)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "compress", "splice"]
    if "skinny" in which:
        bench_skinny()
    if "fused" in which:
        bench_fused()
    if "compress" in which:
        bench_compress()
    if "splice" in which:
        bench_splice()
    if "gemm" in which:
        bench_gemm()
    if "attention" in which:
        bench_attention()
    if "transforms" in which:
        bench_transforms()
