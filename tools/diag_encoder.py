"""GPU diagnostic: per-layer error of the CUDA encoder against the CPU oracle (fp32 and bf16-operand
emulation) on a synthetic recipe.  Usage: python tools/diag_encoder.py tiny|tiny_bn|u2pp_small"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import batch_inputs, err, oracle_cfg  # noqa: E402
from oracle import wenet_oracle as O  # noqa: E402
from wenet_b200 import synth  # noqa: E402
from wenet_b200.asr_model import B200ASRModel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else -1
left = int(sys.argv[3]) if len(sys.argv) > 3 else -1
cfg = synth.recipe(name)
sd = synth.synth_state_dict(cfg)
ns = [32000 + 123, 20800, 48000] if name.startswith("tiny") else [48000, 30000]
pcm, xs, lens = batch_inputs(ns, lambda p: O.fbank(p.float()))
ecfg = oracle_cfg(cfg, sd)
with torch.no_grad():
    taps32, tapsq = [], []
    o32, mask = O.encoder_forward(sd, ecfg, xs, lens, chunk, left, None, taps32)
    oq, _ = O.encoder_forward(sd, ecfg, xs, lens, chunk, left, O.bf16_round, tapsq)
m = B200ASRModel(cfg, sd)
m.keep_layer_dump = True
eo = m._encode(xs.cuda(), lens.cuda(), chunk, left)
torch.cuda.synchronize()
dump = eo.dump.cpu()
el = mask.squeeze(1).sum(1).tolist()
print("enc lens", el, "packed rows", eo.rows)


def packed(t):
    return torch.cat([t[b, :el[b]] for b in range(len(el))], 0)


for i in range(dump.shape[0]):
    a32, aq = packed(taps32[i]), packed(tapsq[i])
    print("tap %2d  vs fp32 oracle max %.3e mean %.3e | vs bf16-emulated oracle max %.3e mean %.3e | oracle fp32-vs-bf16 max %.3e  (rms %.3f)"
          % ((i,) + err(dump[i], a32) + err(dump[i], aq) + (err(a32, aq)[0], float(a32.pow(2).mean().sqrt()))))
out = eo.f32.cpu()
print("final   vs fp32 oracle max %.3e mean %.3e | vs bf16-emulated max %.3e mean %.3e | oracle fp32-vs-bf16 max %.3e mean %.3e"
      % (err(out, packed(o32)) + err(out, packed(oq)) + err(packed(o32), packed(oq))))
