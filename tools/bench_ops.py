"""Per-kernel micro-benchmarks at the bench shapes (S config, 64 x 30 s -> M = 47872 packed frames).
Each op is timed alone with CUDA events, L2 flushed (256 MB memset) before every timed launch.
    python tools/bench_ops.py [filter-substring]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
import ops  # noqa: E402
from wenet_b200 import _lib  # noqa: E402

lib = _lib.load()

flt = sys.argv[1] if len(sys.argv) > 1 else ""
dev = "cuda"
M, d, ff, H, V = 64 * 748, 256, 2048, 4, 4233
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
spin = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(1)


def rb(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


def timeit(name, fn, flops=0.0, bytes_=0.0, iters=8):
    if flt and flt not in name:
        return
    for _ in range(2):
        fn()
    ts = []
    have_diag = lib.wb_gemm_diag(None, 1) == 0     # only in a -DWB_GEMM_DIAG build
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()    # ~45 us of GPU work: the host enqueues the timed launch behind it, so no launch gap is timed
        spin.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    if name.startswith("gemm") and have_diag:
        dg = (C.c_uint64 * 12)()
        lib.wb_gemm_diag(dg, 1)
        n_l = iters
        ctas = 148
        tot = max(dg[6], 1)
        extra_diag = "   [per-CTA share of lifetime: prod-wait %.0f%% mma-wait-data %.0f%% mma-wait-acc %.0f%% epi-wait-acc %.0f%% " \
                     "epi-wait-stage %.0f%% epi-loop %.0f%% | epi per tile: ld-wait %.0f math %.0f store %.0f cyc; %.0f cyc/CTA, %.1f tiles/CTA]" % (
                         100 * dg[0] / tot, 100 * dg[1] / tot, 100 * dg[2] / tot, 100 * dg[3] / tot, 100 * dg[4] / tot,
                         100 * dg[5] / tot, dg[8] / max(dg[7], 1), dg[9] / max(dg[7], 1), dg[10] / max(dg[7], 1),
                         tot / (n_l * ctas), dg[7] / (n_l * ctas))
    else:
        extra_diag = ""
    ts.sort()
    us = ts[len(ts) // 2]
    extra = ""
    if flops:
        extra += "  %.0f TFLOP/s" % (flops / us / 1e6)
    if bytes_:
        extra += "  %.0f GB/s" % (bytes_ / us / 1e3)
    print("%-34s %8.1f us (min %.1f)%s%s" % (name, us, ts[0], extra, extra_diag))


a256 = rb(M, d)
a2048 = rb(M, ff, scale=0.3)
x = torch.randn(M, d, device=dev, generator=g)
for name, N, K, epi, A in [("gemm ffn1 silu N2048 K256", ff, d, ops.EPI_BF16_SILU, a256),
                           ("gemm ffn2 resid N256 K2048", d, ff, ops.EPI_RESID_F32, a2048),
                           ("gemm qkv bf16 N768 K256", 3 * d, d, ops.EPI_BF16, a256),
                           ("gemm out resid N256 K256", d, d, ops.EPI_RESID_F32, a256),
                           ("gemm pw1 glu N512 K256", 2 * d, d, ops.EPI_GLU_BF16, a256),
                           ("gemm ctc f32 N4233 K256", V, d, ops.EPI_F32, a256),
                           ("gemm relu N2048 K256", ff, d, ops.EPI_BF16_RELU, a256)]:
    w = rb(N, K, scale=1.0 / math.sqrt(K))
    b = torch.randn(N, device=dev, generator=g)
    on = N // 2 if epi == ops.EPI_GLU_BF16 else N
    if epi in (ops.EPI_RESID_F32, ops.EPI_F32):
        out = torch.zeros(M, (on + 7) // 8 * 8, device=dev)
    else:
        out = torch.empty(M, on, device=dev, dtype=torch.bfloat16)
    timeit(name, lambda A=A, w=w, b=b, epi=epi, out=out: ops.gemm(A, w, b, epi, 1.0, out=out), flops=2.0 * M * N * K)

gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
for name, K, A in [("gemm+ln out resid N256 K256", d, a256), ("gemm+ln ffn2 resid N256 K2048", ff, a2048)]:
    w = rb(d, K, scale=1.0 / math.sqrt(K))
    b = torch.randn(d, device=dev, generator=g)
    xx = torch.zeros(M, d, device=dev)
    timeit(name, lambda A=A, w=w, b=b, xx=xx: ops.gemm_resid_ln(A, w, b, xx, gamma, beta, alpha=1e-3), flops=2.0 * M * d * K)
timeit("layernorm M x 256 -> bf16", lambda: ops.layernorm(x, gamma, beta), bytes_=M * d * 6.0)

B, T = 64, 748
starts = (torch.arange(B, device=dev, dtype=torch.int32) * T)
lens = torch.full((B,), T, device=dev, dtype=torch.int32)
qkv = rb(M, 3 * d)
kp = rb(M, d)
kb = torch.randn(M, H, device=dev, generator=g)
timeit("attention 64 x 748, 4 heads", lambda: ops.attention(qkv, kp, qkv, starts, lens, starts, lens, H, kbias=kb, q_col0=0,
                                                             k_col0=0, v_col0=2 * d, max_q_len=T),
       flops=4.0 * B * H * T * T * 64)
wdw = torch.randn(d, 8, device=dev, generator=g)
bdw = torch.randn(d, device=dev, generator=g)
timeit("dwconv k8 causal + LN + SiLU", lambda: ops.dwconv(a256, starts, lens, starts, wdw, bdw, gamma, beta, 8, True),
       bytes_=M * d * 4.0)
logits = torch.randn(M, 4240, device=dev, generator=g)
timeit("logsoftmax_topk V4233 k10", lambda: ops.logsoftmax_topk(logits, V, 10), bytes_=M * V * 8.0)
logits2 = torch.randn(M, 4240, device=dev, generator=g)
logits2[:, 0] += 12.0
timeit("lse_topk (no write-back) k10", lambda: ops.lse_topk(logits2, V, 10), bytes_=M * V * 4.0)
