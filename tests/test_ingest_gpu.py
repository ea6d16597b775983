"""wenet_b200.ingest on the GPU: wav files -> pinned int16 -> H2D -> fused fbank -> decode, against decode() fed with the
same utterances directly (results must be identical: the same kernels on the same samples, only the batching differs -
packed rows make every utterance's result independent of its batch neighbours)."""
import wave

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import SEED
from wenet_b200 import synth


def test_transcribe_files_equals_direct_decode(tmp_path):
    from wenet_b200 import ingest
    from wenet_b200.asr_model import B200ASRModel
    from wenet_b200.fbank import FbankExtractor
    cfg = synth.recipe("tiny")
    model = B200ASRModel(cfg, synth.synth_state_dict(cfg, seed=SEED))
    ns = [32000 + 123, 20800, 48000, 16000, 40007, 27000, 52000]
    pcm = synth.synth_pcm(len(ns), ns, seed=SEED)
    paths = []
    for i, n in enumerate(ns):
        p = str(tmp_path / ("u%d.wav" % i))
        with wave.open(p, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(pcm[i, :n].numpy().astype("<i2").tobytes())
        paths.append(p)
    keys = ["u%d" % i for i in range(len(ns))]
    methods = ["ctc_greedy_search", "ctc_prefix_beam_search", "attention_rescoring"]
    got = ingest.transcribe_files(model, paths, keys, methods, beam_size=4, ctc_weight=0.5, reverse_weight=0.3,
                                  max_batch_seconds=8.0, max_batch_size=3)       # forces several ragged batches
    fb = FbankExtractor(80)
    feats = fb(pcm.cuda(), torch.tensor(ns, dtype=torch.int32, device="cuda"))
    lens = torch.tensor([fb.num_frames(n) for n in ns], dtype=torch.int64)
    ref = model.decode(methods, feats[:, :int(lens.max())].contiguous(), lens.cuda(), beam_size=4, ctc_weight=0.5,
                       reverse_weight=0.3)
    for m in methods:
        assert list(got[m].keys()) == keys
        for i, k in enumerate(keys):
            assert list(got[m][k].tokens) == list(ref[m][i].tokens), (m, k)
            if m != "ctc_greedy_search":
                assert [list(h) for h in got[m][k].nbest] == [list(h) for h in ref[m][i].nbest], (m, k)
