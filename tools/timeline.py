"""GPU timeline of the bench workload with N batches in flight (torch.profiler / CUPTI, no nsys in the image).
Prints: wall time per step, union of GPU-busy time, per-stream busy time, the largest idle gaps and what ran
around them.  python tools/timeline.py [inflight] [steps]"""
import json
import os
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wenet_b200 import _lib, synth  # noqa: E402
from wenet_b200.asr_model import B200ASRModel  # noqa: E402
from wenet_b200.fbank import FbankExtractor  # noqa: E402

inflight = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B, n = 64, 480000
cfg = synth.recipe("u2pp_small")
model = B200ASRModel(cfg, synth.synth_state_dict(cfg, seed=777))
fb = FbankExtractor(80)
pcm = synth.synth_pcm(4, n, seed=1).repeat(B // 4, 1).contiguous().cuda()
ns = torch.full((B,), n, dtype=torch.int32, device="cuda")
flens = torch.full((B,), fb.num_frames(n), dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.wb_set_sm_reserve(8 if inflight > 1 else 0)
models = [model] + [model.clone_shared() for _ in range(inflight - 1)]
streams = [torch.cuda.Stream() for _ in range(inflight)]


def step(m):
    feats = fb(pcm, ns)
    return m.decode(["attention_rescoring"], feats, flens, beam_size=10, ctc_weight=0.5, reverse_weight=0.3)


def run(nsteps):
    nxt = [0]
    lock = threading.Lock()

    def worker(w):
        with torch.cuda.stream(streams[w]):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= nsteps:
                    break
                step(models[w])
            streams[w].synchronize()
    ths = [threading.Thread(target=worker, args=(w,)) for w in range(inflight)]
    [t.start() for t in ths]
    [t.join() for t in ths]


run(2 * inflight)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(steps)
    torch.cuda.synchronize()
path = "gpurun_out/timeline_%d.json" % inflight
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
busy, cur_s, cur_e = 0.0, None, None
gaps = []
for e in ev:
    s, f = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, cur_e - t0, e["name"][:50]))
        cur_s, cur_e = s, f
    else:
        cur_e = max(cur_e, f)
busy += cur_e - cur_s
print("span %.2f ms for %d steps = %.2f ms/step; GPU busy (union) %.2f ms = %.1f%%" %
      ((t1 - t0) / 1e3, steps, (t1 - t0) / 1e3 / steps, busy / 1e3, 100 * busy / (t1 - t0)))
per = {}
for e in ev:
    k = e["args"].get("stream", -1)
    per[k] = per.get(k, 0) + e["dur"]
print("per-stream busy ms:", {k: round(v / 1e3, 2) for k, v in per.items()})
print("idle gaps total %.2f ms; top gaps (gap ms, at ms, next kernel):" % (sum(g[0] for g in gaps) / 1e3))
for g in sorted(gaps, reverse=True)[:15]:
    print("  %.3f  @%.2f  %s" % (g[0] / 1e3, g[1] / 1e3, g[2]))
# time where exactly one stream is active vs two
pts = []
for e in ev:
    pts.append((e["ts"], 1))
    pts.append((e["ts"] + e["dur"], -1))
pts.sort()
lvl, last, hist = 0, pts[0][0], {}
for t, d in pts:
    hist[lvl] = hist.get(lvl, 0) + (t - last)
    last = t
    lvl += d
print("concurrency histogram ms:", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
