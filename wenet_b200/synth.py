"""Deterministic synthetic weights and audio (no network, no checkpoints, no torch RNG):
the same `state_dict` can be regenerated bit-for-bit on any box from (configs, seed), so parity tests
on the GPU box load exactly the weights the golden fixtures were produced with in the build
container (oracle/make_goldens.py feeds them to the UNMODIFIED reference).

Key names and shapes are the reference's own; distributions follow torch's
default initialisers closely enough to give well-conditioned activations:
  Linear / Conv weights and biases  ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))
  LayerNorm / BatchNorm             weight 1 + 0.1 U(-1,1), bias 0.1 U(-1,1), var in [0.5, 1.5]
  pos_bias_u / pos_bias_v           xavier uniform
  Embedding                         U(-sqrt(3), sqrt(3))   (unit variance)
The CTC head is sharpened (weight * alpha, blank bias + beta) so that posteriors are peaky like a
trained model's (flat random posteriors make beam search pathological).
"""
import hashlib
import math
from typing import Dict

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    h = hashlib.sha256(("%d:%s" % (seed, name)).encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def _uniform(seed, name, shape, bound) -> torch.Tensor:
    u = _rng(seed, name).random(size=shape)  # float64 in [0, 1)
    return torch.from_numpy(((u * 2.0 - 1.0) * bound).astype(np.float32))


def sinusoid_pe(max_len: int, d: int) -> torch.Tensor:
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def whisper_sinusoids(max_len: int, d: int) -> torch.Tensor:
    """WhisperPositionalEncoding (wenet/models/transformer/embedding.py:150-164)"""
    inc = np.log(10000) / (d // 2 - 1)
    inv = torch.exp(-inc * torch.arange(d // 2))
    st = torch.arange(max_len)[:, np.newaxis] * inv[np.newaxis, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def synth_whisper_state_dict(configs: dict, seed: int = 777, emb_bound: float = 1.5, eos_beta: float = 12.0,
                             pe_bound: float = 0.3) -> Dict[str, torch.Tensor]:
    """Whisper-shaped weights under the reference's key names (TransformerEncoder with Conv1dSubsampling2 /
    WhisperPositionalEncoding, TransformerDecoder with LearnablePositionalEncoding, tied output embedding; key_bias and
    src_key_bias false: no linear_k.bias keys).  The <eot> output bias is raised by eos_beta so that beam search ends
    after a few tens of tokens as it does with a trained model."""
    enc, dec = configs["encoder_conf"], configs["decoder_conf"]
    d, h, ff, L = int(enc["output_size"]), int(enc["attention_heads"]), int(enc["linear_units"]), int(enc["num_blocks"])
    idim, V = int(configs["input_dim"]), int(configs["output_dim"])
    dff, DL = int(dec["linear_units"]), int(dec["num_blocks"])
    st = configs["tokenizer_conf"]["special_tokens"]
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, extra_shape=(), bias=True):
        b = 1.0 / math.sqrt(in_f * int(np.prod(extra_shape)) if extra_shape else in_f)
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (out_f, in_f) + tuple(extra_shape), b)
        if bias:
            sd[name + ".bias"] = _uniform(seed, name + ".bias", (out_f,), b)

    def norm(name):
        sd[name + ".weight"] = 1.0 + _uniform(seed, name + ".weight", (d,), 0.1)
        sd[name + ".bias"] = _uniform(seed, name + ".bias", (d,), 0.1)

    def attn(p, key_bias):
        lin(p + ".linear_q", d, d)
        lin(p + ".linear_k", d, d, bias=key_bias)
        lin(p + ".linear_v", d, d)
        lin(p + ".linear_out", d, d)

    lin("encoder.embed.conv.0", d, idim, (3,))
    lin("encoder.embed.conv.2", d, d, (3,))
    sd["encoder.embed.pos_enc.pe"] = whisper_sinusoids(1500, d).unsqueeze(0)
    norm("encoder.after_norm")
    for i in range(L):
        p = "encoder.encoders.%d" % i
        attn(p + ".self_attn", bool(enc.get("key_bias", True)))
        lin(p + ".feed_forward.w_1", ff, d)
        lin(p + ".feed_forward.w_2", d, ff)
        norm(p + ".norm1")
        norm(p + ".norm2")
    lin("ctc.ctc_lo", V, d)
    sd["decoder.embed.0.weight"] = _uniform(seed, "decoder.emb", (V, d), emb_bound)
    sd["decoder.embed.1.pe"] = _uniform(seed, "decoder.pe", (1, 448, d), pe_bound)
    norm("decoder.after_norm")
    if dec.get("tie_word_embedding", False):
        sd["decoder.output_layer.weight"] = sd["decoder.embed.0.weight"]
    else:
        sd["decoder.output_layer.weight"] = _uniform(seed, "decoder.out", (V, d), emb_bound)
    ob = _uniform(seed, "decoder.output_layer.bias", (V,), 0.1)
    ob[int(st["eot"])] += eos_beta
    sd["decoder.output_layer.bias"] = ob
    for i in range(DL):
        p = "decoder.decoders.%d" % i
        attn(p + ".self_attn", bool(dec.get("key_bias", True)))
        attn(p + ".src_attn", bool(dec.get("src_key_bias", True)))
        lin(p + ".feed_forward.w_1", dff, d)
        lin(p + ".feed_forward.w_2", d, dff)
        for n in ("norm1", "norm2", "norm3"):
            norm(p + "." + n)
    return sd


def synth_whisper_state_dict_fast(configs: dict, seed: int = 777, eos_beta: float = -1e4) -> Dict[str, torch.Tensor]:
    """Benchmark-only variant for Whisper-large-sized models (1.5 B parameters): same keys / shapes / distributions as
    synth_whisper_state_dict but drawn with torch's CPU generator (seconds instead of minutes; not bit-reproducible across
    torch versions, which a throughput run does not need).  eos_beta = -1e4 suppresses <eot>, so attention decoding runs a
    fixed number of steps."""
    tmpl_cfg = dict(configs)
    g = torch.Generator().manual_seed(seed)
    enc, dec = configs["encoder_conf"], configs["decoder_conf"]
    d, ff, L = int(enc["output_size"]), int(enc["linear_units"]), int(enc["num_blocks"])
    idim, V = int(configs["input_dim"]), int(configs["output_dim"])
    dff, DL = int(dec["linear_units"]), int(dec["num_blocks"])
    st = configs["tokenizer_conf"]["special_tokens"]
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound

    def lin(name, out_f, in_f, extra=(), bias=True):
        b = 1.0 / math.sqrt(in_f * int(np.prod(extra)) if extra else in_f)
        sd[name + ".weight"] = uni((out_f, in_f) + tuple(extra), b)
        if bias:
            sd[name + ".bias"] = uni((out_f,), b)

    def norm(name):
        sd[name + ".weight"] = 1.0 + uni((d,), 0.1)
        sd[name + ".bias"] = uni((d,), 0.1)

    def attn(p, key_bias):
        lin(p + ".linear_q", d, d)
        lin(p + ".linear_k", d, d, bias=key_bias)
        lin(p + ".linear_v", d, d)
        lin(p + ".linear_out", d, d)

    lin("encoder.embed.conv.0", d, idim, (3,))
    lin("encoder.embed.conv.2", d, d, (3,))
    sd["encoder.embed.pos_enc.pe"] = whisper_sinusoids(1500, d).unsqueeze(0)
    norm("encoder.after_norm")
    for i in range(L):
        p = "encoder.encoders.%d" % i
        attn(p + ".self_attn", bool(enc.get("key_bias", True)))
        lin(p + ".feed_forward.w_1", ff, d)
        lin(p + ".feed_forward.w_2", d, ff)
        norm(p + ".norm1")
        norm(p + ".norm2")
    sd["decoder.embed.0.weight"] = uni((V, d), 0.5)
    sd["decoder.embed.1.pe"] = uni((1, 448, d), 0.3)
    norm("decoder.after_norm")
    sd["decoder.output_layer.weight"] = sd["decoder.embed.0.weight"] if dec.get("tie_word_embedding", False) else uni((V, d), 0.5)
    ob = uni((V,), 0.1)
    ob[int(st["eot"])] += eos_beta
    sd["decoder.output_layer.bias"] = ob
    for i in range(DL):
        p = "decoder.decoders.%d" % i
        attn(p + ".self_attn", bool(dec.get("key_bias", True)))
        attn(p + ".src_attn", bool(dec.get("src_key_bias", True)))
        lin(p + ".feed_forward.w_1", dff, d)
        lin(p + ".feed_forward.w_2", d, dff)
        for n in ("norm1", "norm2", "norm3"):
            norm(p + "." + n)
    return sd


def synth_state_dict(configs: dict, seed: int = 777, ctc_alpha: float = 8.0, ctc_blank_beta: float = None,
                     with_pe: bool = True) -> Dict[str, torch.Tensor]:
    if configs.get("model") == "whisper":
        return synth_whisper_state_dict(configs, seed)
    enc = configs["encoder_conf"]
    dec = configs.get("decoder_conf", {})
    d = int(enc.get("output_size", 256))
    h = int(enc.get("attention_heads", 4))
    ff = int(enc.get("linear_units", 2048))
    L = int(enc.get("num_blocks", 6))
    K = int(enc.get("cnn_module_kernel", 15))
    bn = enc.get("cnn_module_norm", "batch_norm") == "batch_norm"
    idim = int(configs["input_dim"])
    V = int(configs["output_dim"])
    F2 = ((idim - 1) // 2 - 1) // 2
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, extra_shape=()):
        b = 1.0 / math.sqrt(in_f * int(np.prod(extra_shape)) if extra_shape else in_f)
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (out_f, in_f) + tuple(extra_shape), b)
        sd[name + ".bias"] = _uniform(seed, name + ".bias", (out_f,), b)

    def norm(name, n=d):
        sd[name + ".weight"] = 1.0 + _uniform(seed, name + ".weight", (n,), 0.1)
        sd[name + ".bias"] = _uniform(seed, name + ".bias", (n,), 0.1)

    if configs.get("cmvn", None) is not None:
        # closed-form fit of the per-bin mean / std of synth_pcm()'s log-mel features
        f = torch.arange(idim, dtype=torch.float32)
        sd["encoder.global_cmvn.mean"] = 9.8 + 11.5 * (1.0 - torch.exp(-f / 30.0))
        sd["encoder.global_cmvn.istd"] = torch.full((idim,), 1.0 / 3.0)
    lin("encoder.embed.conv.0", d, 1, (3, 3))
    lin("encoder.embed.conv.2", d, d, (3, 3))
    lin("encoder.embed.out.0", d, d * F2)
    if with_pe:
        sd["encoder.embed.pos_enc.pe"] = sinusoid_pe(5000, d).unsqueeze(0)
    norm("encoder.after_norm")
    for i in range(L):
        p = "encoder.encoders.%d" % i
        xb = math.sqrt(6.0 / (h + d // h))
        sd[p + ".self_attn.pos_bias_u"] = _uniform(seed, p + ".pos_bias_u", (h, d // h), xb)
        sd[p + ".self_attn.pos_bias_v"] = _uniform(seed, p + ".pos_bias_v", (h, d // h), xb)
        for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
            lin(p + ".self_attn." + n, d, d)
        sd[p + ".self_attn.linear_pos.weight"] = _uniform(seed, p + ".linear_pos", (d, d), 1.0 / math.sqrt(d))
        for f in ("feed_forward", "feed_forward_macaron"):
            lin(p + "." + f + ".w_1", ff, d)
            lin(p + "." + f + ".w_2", d, ff)
        lin(p + ".conv_module.pointwise_conv1", 2 * d, d, (1,))
        lin(p + ".conv_module.depthwise_conv", d, 1, (K,))
        norm(p + ".conv_module.norm")
        if bn:
            sd[p + ".conv_module.norm.running_mean"] = _uniform(seed, p + ".bn.mean", (d,), 0.1)
            sd[p + ".conv_module.norm.running_var"] = 1.0 + _uniform(seed, p + ".bn.var", (d,), 0.5)
            sd[p + ".conv_module.norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        lin(p + ".conv_module.pointwise_conv2", d, d, (1,))
        for n in ("norm_ff", "norm_mha", "norm_ff_macaron", "norm_conv", "norm_final"):
            norm(p + "." + n)
    lin("ctc.ctc_lo", V, d)
    if ctc_blank_beta is None:
        # blank logit offset ~ expected maximum of the V-1 non-blank logits (std alpha * 0.58) + 1 sigma:
        # gives a blank fraction of 0.7-0.9 like a trained CTC model
        ctc_blank_beta = ctc_alpha * 0.58 * (math.sqrt(2.0 * math.log(V)) + (1.0 if V > 100 else -0.1))
    sd["ctc.ctc_lo.weight"] = sd["ctc.ctc_lo.weight"] * ctc_alpha
    sd["ctc.ctc_lo.bias"] = sd["ctc.ctc_lo.bias"].clone()
    sd["ctc.ctc_lo.bias"][0] += ctc_blank_beta

    dec_type = configs.get("decoder", "bitransformer")
    dff = int(dec.get("linear_units", 2048))

    def decoder(prefix, n_layers):
        sd[prefix + ".embed.0.weight"] = _uniform(seed, prefix + ".emb", (V, d), math.sqrt(3.0))
        if with_pe:
            sd[prefix + ".embed.1.pe"] = sinusoid_pe(5000, d).unsqueeze(0)
        norm(prefix + ".after_norm")
        lin(prefix + ".output_layer", V, d)
        for i in range(n_layers):
            p = "%s.decoders.%d" % (prefix, i)
            for a in ("self_attn", "src_attn"):
                for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
                    lin(p + "." + a + "." + n, d, d)
            lin(p + ".feed_forward.w_1", dff, d)
            lin(p + ".feed_forward.w_2", d, dff)
            for n in ("norm1", "norm2", "norm3"):
                norm(p + "." + n)

    if dec_type == "bitransformer":
        decoder("decoder.left_decoder", int(dec.get("num_blocks", 6)))
        decoder("decoder.right_decoder", int(dec.get("r_num_blocks", 0)))
    else:
        decoder("decoder", int(dec.get("num_blocks", 6)))
    return sd


def synth_pcm(batch: int, num_samples, seed: int = 777, sigma: float = 1200.0) -> torch.Tensor:
    """int16 PCM (batch, max_n): speech-like NON-stationary synthetic audio — a sequence of 60-320 ms
    segments, each with its own tone pair, amplitude and noise level (silence with probability 0.25) —
    so that encoder outputs and CTC posteriors vary over time like real speech does (stationary noise
    gives time-invariant posteriors, which makes every search degenerate).  Clipped to int16.
    num_samples: int or list of ints; shorter rows are zero padded."""
    ns = [int(num_samples)] * batch if np.isscalar(num_samples) else [int(n) for n in num_samples]
    n_max = (max(ns) + 7) // 8 * 8
    out = np.zeros((batch, n_max), dtype=np.int16)
    for b, n in enumerate(ns):
        r = _rng(seed, "pcm%d" % b)
        u1 = np.maximum(r.random(size=n), 1e-12)
        u2 = r.random(size=n)
        noise = np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)      # Box-Muller: raw uniforms only
        x = np.zeros(n)
        t = np.arange(n) / 16000.0
        pos = 0
        while pos < n:
            seg = int(16000 * (0.06 + 0.26 * r.random()))
            e = min(n, pos + seg)
            kind = r.random()
            f1 = 120.0 + 3000.0 * r.random()
            f2 = 300.0 + 5000.0 * r.random()
            amp = 600.0 + 6000.0 * r.random()
            ns_amp = sigma * (0.2 + 1.5 * r.random())
            if kind < 0.25:
                amp, ns_amp = 0.0, 0.05 * sigma
            tt = t[pos:e]
            x[pos:e] = amp * (np.sin(2 * np.pi * f1 * tt) + 0.6 * np.sin(2 * np.pi * f2 * tt)) + ns_amp * noise[pos:e]
            pos = e
        out[b, :n] = np.clip(np.round(x), -32767, 32767).astype(np.int16)
    return torch.from_numpy(out)


# the recipe configurations of SURVEY.md section 8 (values from the reference's yaml files:
# examples/aishell/s0/conf/train_u2++_conformer.yaml, train_unified_conformer.yaml, train_conformer.yaml,
# examples/wenetspeech/s0/conf/train_u2++_conformer.yaml)
def recipe(name: str) -> dict:
    def enc(d, h, ff, L, K, causal, norm, dyn):
        return dict(output_size=d, attention_heads=h, linear_units=ff, num_blocks=L, dropout_rate=0.1,
                    positional_dropout_rate=0.1, attention_dropout_rate=0.1, input_layer="conv2d",
                    normalize_before=True, cnn_module_kernel=K, use_cnn_module=True, activation_type="swish",
                    pos_enc_layer_type="rel_pos", selfattention_layer_type="rel_selfattn", causal=causal,
                    use_dynamic_chunk=dyn, cnn_module_norm=norm, use_dynamic_left_chunk=False)

    def dec(h, ff, n, r=None):
        c = dict(attention_heads=h, linear_units=ff, num_blocks=n, dropout_rate=0.1, positional_dropout_rate=0.1,
                 self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1)
        if r is not None:
            c["r_num_blocks"] = r
        return c

    # cmvn: the recipes use global CMVN (examples/aishell/s0/conf/*.yaml `cmvn: global_cmvn`); the
    # statistics are part of the state_dict (encoder.global_cmvn.mean / istd)
    base = dict(input_dim=80, cmvn="global_cmvn", cmvn_conf={"cmvn_file": None, "is_json_cmvn": True},
                encoder="conformer", tokenizer="char", tokenizer_conf={}, ctc="ctc",
                ctc_conf={"ctc_blank_id": 0}, model="asr_model")
    if name == "u2pp_small":        # AISHELL-1 U2++ 12L/256d/4h, K=8 causal LN, bitransformer 3+3
        return dict(base, output_dim=4233, encoder_conf=enc(256, 4, 2048, 12, 8, True, "layer_norm", True),
                    decoder="bitransformer", decoder_conf=dec(4, 2048, 3, 3),
                    model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False, reverse_weight=0.3))
    if name == "u2_small":          # train_unified_conformer.yaml: K=15 causal LN, transformer x6
        return dict(base, output_dim=4233, encoder_conf=enc(256, 4, 2048, 12, 15, True, "layer_norm", True),
                    decoder="transformer", decoder_conf=dec(4, 2048, 6),
                    model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False))
    if name == "conformer_small":   # train_conformer.yaml: K=15 symmetric BatchNorm, transformer x6
        return dict(base, output_dim=4233, encoder_conf=enc(256, 4, 2048, 12, 15, False, "batch_norm", False),
                    decoder="transformer", decoder_conf=dec(4, 2048, 6),
                    model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False))
    if name in ("u2pp_large", "u2pp_large12"):   # WenetSpeech U2++ 512d/8h K=15 causal LN, 3+3
        L = 24 if name == "u2pp_large" else 12
        return dict(base, output_dim=5538, encoder_conf=enc(512, 8, 2048, L, 15, True, "layer_norm", True),
                    decoder="bitransformer", decoder_conf=dec(8, 2048, 3, 3),
                    model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False, reverse_weight=0.3))
    if name == "tiny":              # test-sized U2++ (2L/128d/2h, V=37)
        return dict(base, output_dim=37, encoder_conf=enc(128, 2, 256, 2, 8, True, "layer_norm", True),
                    decoder="bitransformer", decoder_conf=dec(2, 256, 2, 1),
                    model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False, reverse_weight=0.3))
    if name == "tiny_bn":           # test-sized non-streaming variant (symmetric K=15, BatchNorm)
        return dict(base, output_dim=37, encoder_conf=enc(128, 2, 256, 2, 15, False, "batch_norm", False),
                    decoder="transformer", decoder_conf=dec(2, 256, 2),
                    model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False))
    if name == "tiny512":           # test-sized wide variant (2L/512d/8h, K=15 causal) — the WenetSpeech geometry
        return dict(base, output_dim=61, encoder_conf=enc(512, 8, 1024, 2, 15, True, "layer_norm", True),
                    decoder="bitransformer", decoder_conf=dec(8, 1024, 1, 1),
                    model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False, reverse_weight=0.3))
    if name == "whisper_wide":      # the large-v3 WIDTHS (d 1280, 20 heads, ff 5120, V 51 866, 128 mel) with 2 + 2 layers: test size
        c = recipe("whisper_large_v3")
        c["encoder_conf"] = dict(c["encoder_conf"], num_blocks=2)
        c["decoder_conf"] = dict(c["decoder_conf"], num_blocks=2, tie_word_embedding=False)
        return c
    if name in ("whisper_tiny", "whisper_large_v3"):
        # examples/aishell/whisper/conf/finetune_whisper_largev3.yaml (32 + 32 L, d 1280, 20 heads, ff 5120, V 51866, 128 mel);
        # whisper_tiny: the same structure at test size
        big = name == "whisper_large_v3"
        d, h, ff, L, V, mel = (1280, 20, 5120, 32, 51866, 128) if big else (128, 2, 256, 2, 120, 32)
        st = dict(eot=50257, no_speech=50363, no_timestamps=50364, sot=50258, sot_prev=50362, timestamp_begin=50365,
                  transcribe=50360, translate=50359) if big else \
            dict(eot=99, sot=100, translate=104, transcribe=105, no_speech=106, no_timestamps=107, sot_prev=108,
                 timestamp_begin=109)
        return dict(
            input_dim=mel, output_dim=V, cmvn=None, cmvn_conf={"cmvn_file": None, "is_json_cmvn": None},
            encoder="transformer",
            encoder_conf=dict(activation_type="gelu", attention_dropout_rate=0.0, attention_heads=h, dropout_rate=0.0,
                              gradient_checkpointing=False, input_layer="conv1d2", key_bias=False, linear_units=ff,
                              normalize_before=True, num_blocks=L, output_size=d, pos_enc_layer_type="abs_pos_whisper",
                              positional_dropout_rate=0.0, static_chunk_size=-1, use_dynamic_chunk=False,
                              use_dynamic_left_chunk=False),
            decoder="transformer",
            decoder_conf=dict(activation_type="gelu", attention_heads=h, dropout_rate=0.0, gradient_checkpointing=False,
                              input_layer="embed_learnable_pe", key_bias=False, src_key_bias=False, linear_units=ff,
                              normalize_before=True, num_blocks=L, positional_dropout_rate=0.0,
                              self_attention_dropout_rate=0.0, src_attention=True, src_attention_dropout_rate=0.0,
                              tie_word_embedding=big, use_output_layer=True),   # tied weights make a random-init decoder repeat its input token
            tokenizer="whisper",
            tokenizer_conf=dict(bpe_path=None, is_multilingual=True, non_lang_syms_path=None, num_languages=100,
                                special_tokens=st, split_with_space=False, symbol_table_path=None),
            ctc="ctc", ctc_conf=dict(ctc_blank_id=st["no_speech"]),
            model="whisper", model_conf=dict(ctc_weight=0.3, length_normalized_loss=False, lsm_weight=0.1))
    raise KeyError(name)
