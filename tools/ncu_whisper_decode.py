"""Profiling harness for the attention-decoding kernels at the Whisper-large geometry with only 2 + 2 layers (fast to build):
    ncu --set full -k regex:dec_cross_attn_part -c 2 python tools/ncu_whisper_decode.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wenet_b200 import synth  # noqa: E402
from wenet_b200.whisper import B200Whisper  # noqa: E402

cfg = synth.recipe("whisper_large_v3")
cfg["encoder_conf"]["num_blocks"] = 2
cfg["decoder_conf"]["num_blocks"] = 2
model = B200Whisper(cfg, synth.synth_whisper_state_dict_fast(cfg, seed=777))
B, T = 32, 3000
model.max_decode_len = int(sys.argv[1]) if len(sys.argv) > 1 else 12
feats = torch.randn(B, T, 128, device="cuda") * 0.3
lens = torch.full((B,), T, dtype=torch.int64, device="cuda")
for _ in range(2):
    out = model.decode(["attention"], feats, lens, beam_size=10)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
out = model.decode(["attention"], feats, lens, beam_size=10)
ev1.record()
torch.cuda.synchronize()
print("decode (2+2 layers, %d steps): %.2f ms" % (model.last_attention_steps, ev0.elapsed_time(ev1)))
