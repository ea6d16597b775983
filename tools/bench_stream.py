"""Streaming latency of encoder.forward_chunk (BASELINE.json configs[3]):
S U2++ Conformer, chunk 16, required_cache_size 64, xs (1, 67, 80), >= 200 consecutive chunks, batch 1.
Prints one JSON line with p50 / p99 per-chunk latency (CUDA events on the launching stream, after warm-up) and the
chunk real-time factor (16 frames x 40 ms = 0.64 s of audio per chunk).

    python tools/bench_stream.py [--chunks 300] [--warmup 20] [--ctc]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wenet_b200 import _lib, synth  # noqa: E402
from wenet_b200.asr_model import B200ASRModel, StreamingSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=280)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--left", type=int, default=4, help="left chunks kept in the attention cache")
    ap.add_argument("--ctc", action="store_true", help="include the CTC head + per-chunk top-k in the timed region")
    ap.add_argument("--graph", action="store_true", help="steady-state chunk step replayed as a CUDA graph (StreamingSession)")
    a = ap.parse_args()
    cfg = synth.recipe("u2pp_small")
    model = B200ASRModel(cfg, synth.synth_state_dict(cfg, seed=777), with_decoder=False)
    lib = _lib.load()
    enc = model.encoder
    chunk = a.chunk
    window = (chunk - 1) * 4 + 7
    stride = 4 * chunk
    total = a.chunks + a.warmup
    g = torch.Generator().manual_seed(777)
    feats = torch.randn(1, stride * total + window, 80, generator=g).cuda()
    att = torch.zeros(0, 0, 0, 0, device="cuda")
    cnn = torch.zeros(0, 0, 0, 0, device="cuda")
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(total)]
    offset = 0
    n0 = lib.wb_launch_count()
    sess = StreamingSession(model, chunk, a.left) if a.graph else None
    for i in range(total):
        xs = feats[:, i * stride:i * stride + window]
        evs[i][0].record()
        if sess is not None:
            y = sess.step(xs)
        else:
            y, att, cnn = enc.forward_chunk(xs, offset, chunk * a.left, att, cnn)
        if a.ctc:
            model.ctc.log_softmax(y)
        evs[i][1].record()
        offset += y.size(1)
        torch.cuda.synchronize()   # one chunk in flight at a time: this is a latency, not a throughput, measurement
    launches = (lib.wb_launch_count() - n0) / total
    ms = np.array([e0.elapsed_time(e1) for e0, e1 in evs[a.warmup:]])
    out = {"metric": "forward_chunk latency (S U2++, chunk %d, cache %d, batch 1)" % (chunk, chunk * a.left),
           "p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)), "mean_ms": float(ms.mean()),
           "chunks": a.chunks, "warmup": a.warmup, "audio_s_per_chunk": chunk * 0.04,
           "chunk_rtf": float(np.percentile(ms, 50)) / 1e3 / (chunk * 0.04), "launches_per_chunk": launches,
           "with_ctc": bool(a.ctc), "cuda_graph": bool(a.graph)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
