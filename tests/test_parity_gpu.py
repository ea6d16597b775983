"""Parity at the BASELINE configurations and at the tolerance BASELINE.json's north_star states.

Goldens (oracle/make_goldens.py, produced by the UNMODIFIED reference in fp32 on seeded synthetic weights / audio):
  u2pp_small_long    12L/256d/4h U2++ (configs[1] model), one 30 s + one 17 s utterance: encoder_out rows [::4],
                     CTC top-10, greedy / 10-best / attention-rescoring results
  u2pp_large_10s     24L/512d/8h U2++ (configs[2] model), one 10 s utterance, rows [::2]
  u2pp_small_stream  configs[3]: forward_chunk chunk 16 / left 4 over 22 chunks of a 14 s utterance, rows [::2],
                     final cnn cache and the final attention cache of the first / last layer

Tolerances (the gates below):
  PRECISE mode (B200ASRModel(precise=True): bf16x3 GEMMs, fp32 attention / depthwise conv):
      encoder_out            max |diff| <= 1e-3 vs the fp32 reference           (north_star: "within 1e-3")
      CTC log-probs          max |diff| <= 1e-3 on the reference's top-10 tokens of every frame
      greedy / n-best / rescoring token ids identical to the reference's ("beam-search token ids bit-exact")
  bf16 mode (throughput mode, bf16 operands / fp32 accumulate):
      encoder_out            inside the reference's OWN bf16-autocast budget (max 5.9e-2, mean 8.2e-3)
      CTC frame arg-max      agrees with the reference on >= 97 % of the frames; the id agreement of greedy /
                             best-beam / rescoring output with the reference is PRINTED (edit distance based) and must
                             be >= 0.9 (greedy) / >= 0.75 (best beam, rescoring: one flipped near-tie frame moves a whole n-best entry) -
                             bf16 operands cannot avoid flips on near-tie frames; precise mode is the exact one
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import SEED, err, load_golden
from wenet_b200 import synth

TOL_PRECISE = 1e-3
BF16_MAX, BF16_MEAN = 5.9e-2, 8.2e-3


def _gpu_fbank(ns):
    from wenet_b200.fbank import FbankExtractor
    pcm = synth.synth_pcm(len(ns), ns, seed=SEED)
    ex = FbankExtractor(80)
    feats = ex(pcm.cuda(), torch.tensor(ns, dtype=torch.int32, device="cuda"))
    lens = torch.tensor([ex.num_frames(n) for n in ns], dtype=torch.int64)
    return feats[:, :int(lens.max())].contiguous(), lens


def _edit_distance(a, b):
    a, b = list(a), list(b)
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def _agreement(hyp, ref):
    return 1.0 - _edit_distance(hyp, ref) / max(len(ref), 1)


_models = {}


def _model(recipe, precise):
    from wenet_b200.asr_model import B200ASRModel
    key = (recipe, precise)
    if key not in _models:
        _models.clear()     # one resident model at a time (the 24-layer recipe is 0.5 GB in precise mode)
        cfg = synth.recipe(recipe)
        sd = synth.synth_state_dict(cfg, seed=SEED)
        _models[key] = (cfg, B200ASRModel(cfg, sd, precise=precise))
    return _models[key]


@pytest.mark.parametrize("precise", [True, False], ids=["precise", "bf16"])
@pytest.mark.parametrize("gname,recipe", [("tiny", "tiny"), ("tiny_bn", "tiny_bn"), ("u2pp_small", "u2pp_small")])
def test_short_goldens(gname, recipe, precise):
    """Full encoder_out (and, for the tiny recipes, the full CTC log-prob matrix) of the round-1 goldens."""
    g = load_golden(gname)
    cfg, model = _model(recipe, precise)
    feats, lens = _gpu_fbank(g["num_samples"].tolist())
    el = g["enc_lens"].tolist()
    out, masks = model.encoder(feats, lens.cuda(), -1, -1)
    assert masks.squeeze(1).sum(1).cpu().tolist() == el
    lp = model.ctc_logprobs(out).cpu()
    for b, n in enumerate(el):
        mx, mn = err(out[b, :n].cpu(), torch.from_numpy(g["enc_out"][b, :n]))
        print("%s[%s] utt %d encoder_out vs fp32 reference: max %.3e mean %.3e" % (gname, "precise" if precise else "bf16", b, mx, mn))
        assert (mx <= TOL_PRECISE) if precise else (mx < BF16_MAX and mn < BF16_MEAN)
        if "ctc_logp" in g and precise:
            mxl, mnl = err(lp[b, :n], torch.from_numpy(g["ctc_logp"][b, :n]))
            print("%s[precise] utt %d ctc log-probs (all tokens) vs fp32 reference: max %.3e mean %.3e" % (gname, b, mxl, mnl))
            assert mxl <= TOL_PRECISE
    if "enc_out_chunk" in g:
        c, l = [int(v) for v in g["chunk"]]
        outc, _ = model.encoder(feats, lens.cuda(), c, l)
        for b, n in enumerate(el):
            mx, mn = err(outc[b, :n].cpu(), torch.from_numpy(g["enc_out_chunk"][b, :n]))
            print("%s[%s] utt %d chunk-masked (%d, %d) encoder_out: max %.3e mean %.3e" % (gname, "precise" if precise else "bf16", b, c, l, mx, mn))
            assert (mx <= TOL_PRECISE) if precise else (mx < BF16_MAX and mn < BF16_MEAN)


@pytest.mark.parametrize("precise", [True, False], ids=["precise", "bf16"])
@pytest.mark.parametrize("gname,recipe", [("u2pp_small_long", "u2pp_small"), ("u2pp_large_10s", "u2pp_large")])
def test_baseline_size_goldens(gname, recipe, precise):
    g = load_golden(gname)
    cfg, model = _model(recipe, precise)
    tag = "%s[%s]" % (gname, "precise" if precise else "bf16")
    ns = g["num_samples"].tolist()
    feats, lens = _gpu_fbank(ns)
    el = g["enc_lens"].tolist()
    stride = int(g["row_stride"])
    beam, cw = int(g["beam"]), float(g["ctc_weight"])
    rw = cfg["model_conf"].get("reverse_weight", 0.0)
    out, masks = model.encoder(feats, lens.cuda(), -1, -1)
    assert masks.squeeze(1).sum(1).cpu().tolist() == el
    lp = model.ctc_logprobs(out).cpu()
    res = model.decode(["ctc_greedy_search", "ctc_prefix_beam_search", "attention_rescoring"], feats, lens.cuda(),
                       beam_size=beam, ctc_weight=cw, reverse_weight=rw)
    for b, n in enumerate(el):
        mx, mn = err(out[b, :n:stride].cpu(), torch.from_numpy(g["enc_rows%d" % b]))
        ti = torch.from_numpy(g["topk_idx%d" % b].astype(np.int64))
        tv = torch.from_numpy(g["topk_val%d" % b])
        mine = lp[b, :n].gather(1, ti)
        mxl, mnl = err(mine, tv)
        frame_agree = float((lp[b, :n].argmax(-1) == ti[:, 0]).float().mean())
        gr = res["ctc_greedy_search"][b].tokens
        pb = res["ctc_prefix_beam_search"][b]
        ar = res["attention_rescoring"][b]
        n_ref = int(g["nbest_n%d" % b])
        ref_nbest = [g["nbest%d_%d" % (b, i)].tolist() for i in range(n_ref)]
        a_greedy = _agreement(gr, g["greedy%d" % b].tolist())
        a_beam = _agreement(pb.tokens, ref_nbest[0])
        a_resc = _agreement(ar.tokens, g["resc_tokens%d" % b].tolist())
        nbest_same = sum(int(list(h) == r) for h, r in zip(pb.nbest, ref_nbest)) / float(n_ref)
        print("%s utt %d (T'=%d): encoder_out max %.3e mean %.3e | top-%d log-probs max %.3e mean %.3e | frame arg-max "
              "agreement %.4f | id agreement with the reference: greedy %.4f best-beam %.4f rescoring %.4f, n-best lists "
              "identical %.2f" % (tag, b, n, mx, mn, beam, mxl, mnl, frame_agree, a_greedy, a_beam, a_resc, nbest_same))
        if precise:
            assert mx <= TOL_PRECISE, (tag, b, mx)
            assert mxl <= TOL_PRECISE, (tag, b, mxl)
            assert gr == g["greedy%d" % b].tolist()
            assert [list(h) for h in pb.nbest] == ref_nbest
            assert pb.nbest_times == [g["nbest_time%d_%d" % (b, i)].tolist() for i in range(n_ref)]
            assert np.allclose(pb.nbest_scores, g["nbest_scores%d" % b], rtol=0, atol=2e-2)
            # the rescoring decoder stays bf16 in precise mode: the chosen hypothesis may only differ on a near tie
            assert a_resc >= 0.95
        else:
            assert mx < BF16_MAX and mn < BF16_MEAN, (tag, b, mx, mn)
            assert frame_agree >= 0.97, (tag, b, frame_agree)
            assert a_greedy >= 0.9 and min(a_beam, a_resc) >= 0.75, (tag, b, a_greedy, a_beam, a_resc)


@pytest.mark.parametrize("precise", [False, True], ids=["bf16", "precise"])
def test_forward_chunk_16_4_on_12_layers(precise):
    """BASELINE configs[3] geometry: chunk 16 / 4 left chunks on the 12-layer recipe, 22 chunks incl. a ragged last one,
    through encoder.forward_chunk_by_chunk, forward_chunk (caches) and the CUDA-graph StreamingSession.  Precise mode:
    outputs and caches within 1e-3 of the fp32 reference; bf16 mode: inside the reference's bf16 budget."""
    from wenet_b200.asr_model import StreamingSession
    g = load_golden("u2pp_small_stream")
    cfg, model = _model("u2pp_small", precise)
    MAXE, MEANE = (TOL_PRECISE, TOL_PRECISE) if precise else (BF16_MAX, BF16_MEAN)
    c, l = [int(v) for v in g["chunk"]]
    stride = int(g["row_stride"])
    feats, lens = _gpu_fbank([int(g["num_samples"])])
    n0 = int(lens[0])
    xs = feats[0:1, :n0]
    ys, masks = model.encoder.forward_chunk_by_chunk(xs, c, l)
    assert ys.size(1) == int(g["n_out"]) and masks.shape == (1, 1, ys.size(1))
    mx, mn = err(ys[0, ::stride].cpu(), torch.from_numpy(g["stream_rows"]))
    print("u2pp_small[%s] chunk-by-chunk (16, 4) vs fp32 reference: max %.3e mean %.3e over %d chunks"
          % ("precise" if precise else "bf16", mx, mn, int(g["n_chunks"])))
    assert mx <= MAXE and mn <= MEANE
    # explicit forward_chunk loop: caches after the last chunk
    win, hop = (c - 1) * 4 + 7, 4 * c
    att = torch.zeros(0, 0, 0, 0, device="cuda")
    cnn = torch.zeros(0, 0, 0, 0, device="cuda")
    sess = StreamingSession(model, c, l)
    off, outs, souts = 0, [], []
    for cur in range(0, n0 - 7 + 1, hop):
        w = xs[:, cur:min(cur + win, n0)]
        y, att, cnn = model.encoder.forward_chunk(w, off, c * l, att, cnn)
        souts.append(sess.step(w).clone())
        outs.append(y)
        off += y.size(1)
    assert len(outs) == int(g["n_chunks"])
    assert torch.equal(torch.cat(outs, 1), ys)
    assert torch.equal(torch.cat(souts, 1), ys), "graph-replayed session must equal forward_chunk bit for bit"
    assert sess.graph is not None
    layers = g["att_layers"].tolist()
    mxa, mna = err(att[layers].cpu(), torch.from_numpy(g["att_last"]))
    mxc, mnc = err(cnn.cpu(), torch.from_numpy(g["cnn_last"]))
    print("final caches vs reference: att (layers %s) max %.3e mean %.3e | cnn max %.3e mean %.3e" % (layers, mxa, mna, mxc, mnc))
    assert tuple(att.shape[1:]) == tuple(g["att_last"].shape[1:]) and tuple(cnn.shape) == tuple(g["cnn_last"].shape)
    if precise:
        assert mxa <= TOL_PRECISE and mxc <= TOL_PRECISE
    else:
        assert mxa < 2 * BF16_MAX and mna < BF16_MEAN and mxc < 2 * BF16_MAX and mnc < BF16_MEAN
    # streaming == chunk-masked full forward on the GPU (one attention kernel serves both modes)
    full, _ = model.encoder(xs, lens[0:1].cuda(), c, l)
    mx, mn = err(ys.cpu(), full[:, :ys.size(1)].cpu())
    print("chunk-by-chunk vs chunk-masked forward (both CUDA): max %.3e mean %.3e" % (mx, mn))
    assert mx <= MAXE and mn <= MEANE


def test_encoder_default_arguments():
    """encoder(xs, lens) with the reference's default decoding_chunk_size=0 must work for models that do not
    use chunk masks, use static_chunk_size when set, and raise only for use_dynamic_chunk (training-time random chunk)."""
    from wenet_b200.asr_model import B200ASRModel
    cfg = synth.recipe("tiny_bn")                       # no dynamic chunk, static_chunk_size 0
    sd = synth.synth_state_dict(cfg, seed=SEED)
    model = B200ASRModel(cfg, sd, with_decoder=False)
    feats, lens = _gpu_fbank([32000, 20800])
    a, _ = model.encoder(feats, lens.cuda())
    b, _ = model.encoder(feats, lens.cuda(), -1, -1)
    assert torch.equal(a, b)
    cfg2 = synth.recipe("tiny")
    cfg2["encoder_conf"] = dict(cfg2["encoder_conf"], use_dynamic_chunk=False, static_chunk_size=4)
    m2 = B200ASRModel(cfg2, synth.synth_state_dict(cfg2, seed=SEED), with_decoder=False)
    a, _ = m2.encoder(feats, lens.cuda())
    b, _ = m2.encoder(feats, lens.cuda(), 4, -1)
    assert torch.equal(a, b)
    m3 = B200ASRModel(synth.recipe("tiny"), synth.synth_state_dict(synth.recipe("tiny"), seed=SEED), with_decoder=False)
    with pytest.raises(NotImplementedError):
        m3.encoder(feats, lens.cuda())
