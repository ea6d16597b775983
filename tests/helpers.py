"""Shared test helpers: deterministic inputs, golden loading, oracle runs."""
import os

import numpy as np
import torch

from wenet_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 777


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def batch_inputs(ns, feats_fn):
    """synthetic PCM -> per-utterance features (feats_fn(pcm_row_int16[:n]) -> (m, 80)) -> padded batch"""
    pcm = synth.synth_pcm(len(ns), ns, seed=SEED)
    feats = [feats_fn(pcm[b, :n]) for b, n in enumerate(ns)]
    lens = torch.tensor([f.shape[0] for f in feats], dtype=torch.int64)
    T = int(lens.max())
    xs = torch.zeros(len(ns), T, 80)
    for b, f in enumerate(feats):
        xs[b, :f.shape[0]] = f
    return pcm, xs, lens


def oracle_cfg(cfg, sd):
    from oracle import wenet_oracle as O
    e = cfg["encoder_conf"]
    return O.encoder_cfg(sd, e["attention_heads"], e["causal"], e["cnn_module_norm"])


def decoder_cfg(cfg):
    d = cfg["decoder_conf"]
    return dict(bidirectional=cfg["decoder"] == "bitransformer", layers=d["num_blocks"],
                r_layers=d.get("r_num_blocks", 0), heads=d["attention_heads"])


def err(a, b):
    d = (a.float() - b.float()).abs()
    return float(d.max()), float(d.mean())


def fbank_f64(pcm_i16_row, num_mel=80, frame_len=400, shift=160, nfft=512, preemph=0.97):
    """Kaldi fbank (torchaudio/compliance/kaldi.py:514-645, dither 0, povey window) evaluated in float64 on the same
    float32 window / mel-filter constants the reference uses: the "exact arithmetic" yard-stick for fp32 FFT front-ends
    (two fp32 FFTs of a signal with > 70 dB of dynamic range agree only to ~1e-3 in the log-mel domain on weak bins)."""
    from wenet_b200.fbank import _mel_banks, _povey_window
    x = pcm_i16_row.to(torch.float64).numpy()
    n = x.shape[0]
    m = 1 + (n - frame_len) // shift if n >= frame_len else 0
    if m == 0:
        return torch.zeros(0, num_mel, dtype=torch.float64)
    idx = np.arange(frame_len)[None, :] + shift * np.arange(m)[:, None]
    fr = x[idx]
    fr = fr - fr.mean(axis=1, keepdims=True)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    fr = (fr - preemph * prev) * _povey_window(frame_len).to(torch.float64).numpy()[None, :]
    spec = np.abs(np.fft.rfft(fr, n=nfft, axis=1)) ** 2
    mel = _mel_banks(num_mel, nfft, 16000.0).to(torch.float64).numpy()
    e = spec @ mel.T
    return torch.from_numpy(np.log(np.maximum(e, np.finfo(np.float32).eps)))
