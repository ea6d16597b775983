"""One profiled step of the bench workload for Nsight Compute:
    ncu --profile-from-start off ... python tools/ncu_step.py [batch] [seconds]
A warm-up step runs outside the profiled range; cudaProfilerStart/Stop bracket exactly one step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wenet_b200 import synth  # noqa: E402
from wenet_b200.asr_model import B200ASRModel  # noqa: E402
from wenet_b200.fbank import FbankExtractor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
mode = sys.argv[3] if len(sys.argv) > 3 else "attention_rescoring"
n = int(secs * 16000)
cfg = synth.recipe("u2pp_small")
model = B200ASRModel(cfg, synth.synth_state_dict(cfg, seed=777))
fb = FbankExtractor(80)
pcm = synth.synth_pcm(4, n, seed=1).repeat((B + 3) // 4, 1)[:B].contiguous().cuda()
ns = torch.full((B,), n, dtype=torch.int32, device="cuda")
flens = torch.full((B,), fb.num_frames(n), dtype=torch.int64, device="cuda")


def step():
    feats = fb(pcm, ns)
    return model.decode([mode], feats, flens, beam_size=10, ctc_weight=0.5, reverse_weight=0.3)


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
