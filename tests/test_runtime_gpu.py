"""The C++ runtime back-end (runtime/b200_asr_model.h): `B200AsrModel` with the reference
runtime's AsrModel interface, driven the way runtime/core/decoder/asr_decoder.cc drives it (num_frames_for_chunk ->
ForwardEncoder per chunk -> AttentionRescoring), against the Python path on the same model and features: same kernels,
same windows -> CTC log-probs equal, rescoring scores equal to fp32 summation noise."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import SEED
from wenet_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "runtime", "b200_asr_main")


@pytest.mark.parametrize("chunk,left", [(4, 2), (16, -1)])
def test_cpp_asr_model_matches_python_path(tmp_path, chunk, left):
    from wenet_b200.asr_model import B200ASRModel
    from wenet_b200.export import export_model
    from wenet_b200.fbank import FbankExtractor
    if not os.path.exists(EXE):
        from wenet_b200 import build
        build.build_runtime()
    cfg = synth.recipe("tiny")
    sd = synth.synth_state_dict(cfg, seed=SEED)
    wbm = str(tmp_path / "tiny.wbm")
    assert export_model(cfg, sd, wbm) > 50
    model = B200ASRModel(cfg, sd)
    n = 48000 + 777
    pcm = synth.synth_pcm(1, n, seed=SEED)
    fb = FbankExtractor(80)
    feats = fb(pcm.cuda(), torch.tensor([n], dtype=torch.int32, device="cuda"))[:, :fb.num_frames(n)].contiguous()
    fpath, opath, hpath = str(tmp_path / "feats.bin"), str(tmp_path / "out.bin"), str(tmp_path / "hyps.txt")
    f32 = feats[0].cpu().numpy().astype("<f4")
    with open(fpath, "wb") as f:
        f.write(struct.pack("<2i", f32.shape[0], f32.shape[1]))
        f.write(f32.tobytes())
    hyps = [[3, 5, 7, 9], [4, 6], [], [11, 12, 13, 14, 15, 16]]
    with open(hpath, "w") as f:
        for h in hyps:
            f.write(" ".join(str(t) for t in h) + "\n")
    rw = 0.3
    r = subprocess.run([EXE, wbm, fpath, opath, str(chunk), str(left), str(rw), hpath], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = open(opath, "rb").read()
    rows, V = struct.unpack_from("<2i", raw, 0)
    probs = np.frombuffer(raw, dtype="<f4", count=rows * V, offset=8).reshape(rows, V)
    nh, = struct.unpack_from("<i", raw, 8 + rows * V * 4)
    scores = np.frombuffer(raw, dtype="<f4", count=nh, offset=12 + rows * V * 4)
    # Python path: the same windows through encoder.forward_chunk_by_chunk, then ctc / decoder
    ys, _ = model.encoder.forward_chunk_by_chunk(feats, chunk, left)
    lp = model.ctc_logprobs(ys)[0].cpu().numpy()
    assert probs.shape == lp.shape, (probs.shape, lp.shape)
    assert np.abs(probs - lp).max() < 1e-5, np.abs(probs - lp).max()
    L = max(len(h) for h in hyps) + 1
    hp = torch.full((len(hyps), L), model.eos, dtype=torch.long)
    hp[:, 0] = model.sos
    for i, h in enumerate(hyps):
        hp[i, 1:1 + len(h)] = torch.tensor(h, dtype=torch.long)
    hl = torch.tensor([len(h) + 1 for h in hyps])
    out, r_out = model.forward_attention_decoder(hp.cuda(), hl.cuda(), ys, rw)
    out, r_out = out.cpu(), r_out.cpu()
    assert nh == len(hyps)
    for i, h in enumerate(hyps):
        s = sum(float(out[i, j, t]) for j, t in enumerate(h)) + float(out[i, len(h), model.eos])
        rs = sum(float(r_out[i, j, t]) for j, t in enumerate(h[::-1])) + float(r_out[i, len(h), model.eos])
        want = s * (1 - rw) + rs * rw
        assert abs(scores[i] - want) < 1e-3 * max(1.0, abs(want)), (i, scores[i], want)
