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
