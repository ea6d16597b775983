"""Streaming encoder (forward_chunk / forward_chunk_by_chunk) on the GPU against the reference goldens
and the oracle (encoder.py:204-362)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import SEED, batch_inputs, err, load_golden, oracle_cfg
from oracle import wenet_oracle as O
from wenet_b200 import synth


def test_forward_chunk_vs_reference_golden():
    from wenet_b200.asr_model import B200ASRModel
    g = load_golden("tiny")
    cfg = synth.recipe("tiny")
    sd = synth.synth_state_dict(cfg, seed=SEED)
    model = B200ASRModel(cfg, sd)
    ns = g["num_samples"].tolist()
    _, xs, lens = batch_inputs(ns, lambda p: O.fbank(p.float()))
    xs = xs.cuda()
    c, l = [int(v) for v in g["chunk"]]
    win = (c - 1) * 4 + 7
    att = torch.zeros(0, 0, 0, 0, device="cuda")
    cnn = torch.zeros(0, 0, 0, 0, device="cuda")
    y, att, cnn = model.encoder.forward_chunk(xs[0:1, :win], 0, c * l, att, cnn)
    assert tuple(y.shape) == tuple(g["stream_y1"].shape)
    assert tuple(att.shape) == tuple(g["stream_att1"].shape) and tuple(cnn.shape) == tuple(g["stream_cnn1"].shape)
    print("chunk1 y", err(y.cpu(), torch.from_numpy(g["stream_y1"])), "att", err(att.cpu(), torch.from_numpy(g["stream_att1"])),
          "cnn", err(cnn.cpu(), torch.from_numpy(g["stream_cnn1"])))
    assert err(y.cpu(), torch.from_numpy(g["stream_y1"]))[0] < 5.9e-2
    y2, att2, cnn2 = model.encoder.forward_chunk(xs[0:1, 4 * c:4 * c + win], y.size(1), c * l, att, cnn)
    assert tuple(att2.shape) == tuple(g["stream_att2"].shape)
    for got, key in ((y2, "stream_y2"), (att2, "stream_att2"), (cnn2, "stream_cnn2")):
        mx, mn = err(got.cpu(), torch.from_numpy(g[key]))
        print(key, mx, mn)
        assert mx < 5.9e-2 and mn < 8.2e-3
    # whole utterance chunk by chunk == reference forward_chunk_by_chunk; and == chunk-masked full forward
    n0 = int(lens[0])
    ys, masks = model.encoder.forward_chunk_by_chunk(xs[0:1, :n0], c, l)
    mx, mn = err(ys.cpu(), torch.from_numpy(g["stream_out"]))
    print("chunk-by-chunk vs reference", mx, mn)
    assert ys.shape == g["stream_out"].shape and mx < 5.9e-2 and mn < 8.2e-3
    full, _ = model.encoder(xs[0:1, :n0], lens[0:1].cuda(), c, l)
    mx, mn = err(ys.cpu(), full[:, :ys.size(1)].cpu())
    print("chunk-by-chunk vs chunk-masked forward (both CUDA)", mx, mn)
    assert mx < 5.9e-2 and mn < 8.2e-3
    # simulate_streaming decode path runs
    res = model.decode(["ctc_greedy_search"], xs[0:1, :n0], lens[0:1].cuda(), decoding_chunk_size=c,
                       num_decoding_left_chunks=l, simulate_streaming=True)
    assert len(res["ctc_greedy_search"]) == 1


def test_streaming_session_graph_replay_is_bit_identical():
    """StreamingSession replays the steady-state chunk step as a CUDA graph (wb_encoder_forward_chunk_static: position
    offset read on the device, caches rotated by captured copies); outputs must equal forward_chunk() bit for bit."""
    from wenet_b200.asr_model import B200ASRModel, StreamingSession
    cfg = synth.recipe("tiny")
    sd = synth.synth_state_dict(cfg, seed=SEED)
    model = B200ASRModel(cfg, sd, with_decoder=False)
    chunk, left = 4, 2
    window, stride = (chunk - 1) * 4 + 7, 4 * chunk
    n_chunks = 12
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(1, stride * n_chunks + window, 80, generator=g).cuda()
    att = torch.zeros(0, 0, 0, 0, device="cuda")
    cnn = torch.zeros(0, 0, 0, 0, device="cuda")
    sess = StreamingSession(model, chunk, left)
    offset = 0
    replayed = 0
    for i in range(n_chunks):
        xs = feats[:, i * stride:i * stride + window]
        y_ref, att, cnn = model.encoder.forward_chunk(xs, offset, chunk * left, att, cnn)
        offset += y_ref.size(1)
        y = sess.step(xs).clone()
        replayed += int(sess.graph is not None)
        assert torch.equal(y, y_ref), (i, (y - y_ref).abs().max().item())
    assert replayed >= n_chunks - left - 1      # everything after the cache filled up went through the graph
    assert torch.equal(sess.s_att, att) and torch.equal(sess.s_cnn, cnn)


def test_batched_sessions_equal_single_sessions():
    """wb_encoder_forward_chunk_batch: S sessions in lockstep == S independent forward_chunk streams, bit for bit (per-row
    arithmetic does not depend on how many rows a GEMM sees), including the CUDA-graph steady state."""
    from wenet_b200.asr_model import B200ASRModel, BatchedStreamingSessions
    cfg = synth.recipe("u2pp_small")
    sd = synth.synth_state_dict(cfg, seed=SEED)
    model = B200ASRModel(cfg, sd)
    S, chunk, left, n_chunks = 3, 16, 2, 6
    window, hop = (chunk - 1) * 4 + 7, 4 * chunk
    g = torch.Generator().manual_seed(5)
    feats = (torch.randn(S, hop * n_chunks + window, 80, generator=g) * 2 + 10).cuda()
    # reference: one stream at a time through forward_chunk
    singles = []
    for s in range(S):
        att = torch.zeros(0, 0, 0, 0, device="cuda")
        cnn = torch.zeros(0, 0, 0, 0, device="cuda")
        off, ys = 0, []
        for i in range(n_chunks):
            y, att, cnn = model.encoder.forward_chunk(feats[s:s + 1, i * hop:i * hop + window], off, chunk * left, att, cnn)
            off += y.size(1)
            ys.append(y)
        singles.append((torch.cat(ys, 1), att, cnn))
    for use_graph in (False, True):
        sess = BatchedStreamingSessions(model, S, chunk, left, use_graph=use_graph)
        ys = [sess.step(feats[:, i * hop:i * hop + window]).clone() for i in range(n_chunks)]
        yb = torch.cat(ys, 1)
        att_b = sess.s_att if sess.graph is not None else sess.att
        cnn_b = sess.s_cnn if sess.graph is not None else sess.cnn
        assert (sess.graph is not None) == use_graph
        for s in range(S):
            assert torch.equal(yb[s], singles[s][0][0]), (use_graph, s, float((yb[s] - singles[s][0][0]).abs().max()))
            assert torch.equal(att_b[s], singles[s][1])
            assert torch.equal(cnn_b[s], singles[s][2].reshape(cnn_b[s].shape))
