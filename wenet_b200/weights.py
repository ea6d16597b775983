"""Weight packer: reference `state_dict` (key names of SURVEY.md section 8a) -> the tensors
libwenet_b200.so expects (include/wenet_b200.h, wb_model_set_tensor).

All transformations are layout-only or exact algebra on weights:
  * q/k/v (and cross-attention k/v) projections fused into one [3d, d] ([2d, d]) GEMM
  * Conv2d(d, d, 3, 2) weight -> im2col order (kh, kw, c_in); embed Linear columns permuted from the
    reference's (c, f) flattening (subsampling.py:225) to the channels-last (f, c) this build uses
  * pointwise_conv1 rows interleaved [16 value | 16 gate] so GLU is a GEMM epilogue
  * eval BatchNorm folded to scale/shift; GLU(pointwise_conv1(0)) precomputed (see convmod.cu)
  * linear_pos weight split into bf16 [hi | hi | lo] for the one-off bf16x3 position projection
GEMM weights are stored as bf16 (the operand type of tcgen05.mma kind::f16); everything else fp32.
"""
import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _lib
from ._lib import WbModelConfig, check, cur_stream

WB_F32, WB_BF16, WB_I32 = 0, 1, 2


def interleave_glu(w: torch.Tensor, b: torch.Tensor):
    """[value rows (d) | gate rows (d)] -> blocks of 32 rows = 16 value rows + their 16 gate rows."""
    two_d = w.shape[0]
    d = two_d // 2
    assert d % 16 == 0
    idx = []
    for g in range(d // 16):
        idx += list(range(16 * g, 16 * g + 16)) + list(range(d + 16 * g, d + 16 * g + 16))
    idx = torch.tensor(idx, dtype=torch.long, device=w.device)
    return w.index_select(0, idx), b.index_select(0, idx)


def split3_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 [N, K] -> bf16 [N, 3K] = [hi | hi | lo]; pairs with activations packed [hi | lo | hi]."""
    w = w.float()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], dim=1).contiguous()


def sinusoid_pe(max_len: int, d: int) -> torch.Tensor:
    """wenet/models/transformer/embedding.py:50-59"""
    import math
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


class ModelSpec:
    """The subset of the reference's train.yaml this build understands; anything else raises at
    construction (no silent fallbacks — SURVEY.md section 8b 'unsupported-config policy')."""

    def __init__(self, configs: dict):
        enc = dict(configs.get("encoder_conf", {}))
        self.arch = 0            # 0 Conformer, 1 Whisper (TransformerEncoder conv1d2 / abs_pos_whisper / gelu)
        self.dec_flavor = 0      # 0 wenet "embed" + relu, 1 Whisper "embed_learnable_pe" + gelu
        self.dec_max_len = 0
        if configs.get("encoder", "conformer") == "transformer":
            self._init_whisper(configs, enc)
            return
        if configs.get("encoder", "conformer") != "conformer":
            raise NotImplementedError("encoder '%s' is outside the implemented set (conformer, whisper-style transformer)"
                                      % configs.get("encoder"))

        def need(key, allowed, default):
            v = enc.get(key, default)
            if v not in allowed:
                raise NotImplementedError("encoder_conf.%s=%r is outside the implemented set %r" % (key, v, allowed))
            return v

        need("input_layer", ("conv2d",), "conv2d")
        need("pos_enc_layer_type", ("rel_pos",), "rel_pos")
        need("selfattention_layer_type", ("rel_selfattn",), "rel_selfattn")
        need("activation_type", ("swish",), "swish")
        need("normalize_before", (True,), True)
        need("use_cnn_module", (True,), True)
        need("macaron_style", (True,), True)
        need("layer_norm_type", ("layer_norm",), "layer_norm")
        need("mlp_type", ("position_wise_feed_forward",), "position_wise_feed_forward")
        for k in ("n_kv_head", "head_dim"):
            if enc.get(k) is not None:
                raise NotImplementedError("encoder_conf.%s is not supported" % k)
        if enc.get("use_sdpa", False):
            pass  # numerically the same attention; the flag only selects a torch code path
        self.input_dim = int(configs["input_dim"])
        self.vocab = int(configs["output_dim"])
        self.d_model = int(enc.get("output_size", 256))
        self.heads = int(enc.get("attention_heads", 4))
        self.ffn_dim = int(enc.get("linear_units", 2048))
        self.enc_layers = int(enc.get("num_blocks", 6))
        self.cnn_kernel = int(enc.get("cnn_module_kernel", 15))
        self.cnn_causal = bool(enc.get("causal", False))
        self.cnn_norm = need("cnn_module_norm", ("batch_norm", "layer_norm"), "batch_norm")
        self.use_dynamic_chunk = bool(enc.get("use_dynamic_chunk", False))
        self.static_chunk_size = int(enc.get("static_chunk_size", 0))
        self.ln_eps = float(enc.get("norm_eps", 1e-5))
        if self.d_model != self.heads * 64:
            raise NotImplementedError("attention head size must be 64 (d_model=%d heads=%d)" % (self.d_model, self.heads))
        dec_type = configs.get("decoder", "bitransformer")
        dec = dict(configs.get("decoder_conf", {}))
        if dec_type not in ("transformer", "bitransformer"):
            raise NotImplementedError("decoder '%s' is outside the implemented set" % dec_type)
        self.bidirectional = dec_type == "bitransformer"
        self.dec_layers = int(dec.get("num_blocks", 6))
        self.rdec_layers = int(dec.get("r_num_blocks", 0)) if self.bidirectional else 0
        self.dec_heads = int(dec.get("attention_heads", 4))
        self.dec_ffn_dim = int(dec.get("linear_units", 2048))
        self.dec_ln_eps = float(dec.get("norm_eps", 1e-5))      # decoder.py:83, independent of the encoder's
        if dec.get("activation_type", "relu") != "relu" or dec.get("input_layer", "embed") != "embed" \
                or not dec.get("normalize_before", True) or dec.get("tie_word_embedding", False):
            raise NotImplementedError("decoder_conf outside the implemented set (relu / embed / pre-norm / untied)")
        if self.vocab > 0 and self.d_model != self.dec_heads * 64:      # (vocab 0: encoder-only handle, no decoder at all)
            raise NotImplementedError("decoder head size must be 64")
        st = (configs.get("tokenizer_conf") or {}).get("special_tokens") or {}
        self.sos = int(st.get("<sos>", self.vocab - 1))         # asr_model.py:60-63
        self.eos = int(st.get("<eos>", self.vocab - 1))
        mc = dict(configs.get("model_conf", {}))
        self.reverse_weight = float(mc.get("reverse_weight", 0.0))
        self.ctc_weight = float(mc.get("ctc_weight", 0.5))
        self.max_pos = 5000
        self.has_cmvn = configs.get("cmvn", None) is not None

    def _init_whisper(self, configs: dict, enc: dict):
        """encoder: transformer in the Whisper configuration (examples/aishell/whisper/conf/finetune_whisper_largev3.yaml):
        conv1d2 / abs_pos_whisper / gelu / key_bias false / pre-norm; decoder: transformer with embed_learnable_pe / gelu."""
        def need(conf, key, allowed, default, what):
            v = conf.get(key, default)
            if v not in allowed:
                raise NotImplementedError("%s.%s=%r is outside the implemented set %r" % (what, key, v, allowed))
            return v

        need(enc, "input_layer", ("conv1d2",), "conv2d", "encoder_conf")
        need(enc, "pos_enc_layer_type", ("abs_pos_whisper",), "abs_pos", "encoder_conf")
        need(enc, "activation_type", ("gelu",), "relu", "encoder_conf")
        need(enc, "normalize_before", (True,), True, "encoder_conf")
        need(enc, "selfattention_layer_type", ("selfattn",), "selfattn", "encoder_conf")
        need(enc, "layer_norm_type", ("layer_norm",), "layer_norm", "encoder_conf")
        need(enc, "mlp_type", ("position_wise_feed_forward",), "position_wise_feed_forward", "encoder_conf")
        if enc.get("use_dynamic_chunk", False) or int(enc.get("static_chunk_size", 0)) > 0:
            raise NotImplementedError("chunk masks are not part of the Whisper path")
        for k in ("n_kv_head", "head_dim"):
            if enc.get(k) is not None:
                raise NotImplementedError("encoder_conf.%s is not supported" % k)
        self.arch = 1
        self.input_dim = int(configs["input_dim"])
        self.vocab = int(configs["output_dim"])
        self.d_model = int(enc.get("output_size", 256))
        self.heads = int(enc.get("attention_heads", 4))
        self.ffn_dim = int(enc.get("linear_units", 2048))
        self.enc_layers = int(enc.get("num_blocks", 6))
        self.cnn_kernel, self.cnn_causal, self.cnn_norm = 1, False, "layer_norm"
        self.use_dynamic_chunk, self.static_chunk_size = False, 0
        self.ln_eps = float(enc.get("norm_eps", 1e-5))
        if self.d_model != self.heads * 64:
            raise NotImplementedError("attention head size must be 64 (d_model=%d heads=%d)" % (self.d_model, self.heads))
        if configs.get("decoder", "transformer") != "transformer":
            raise NotImplementedError("decoder '%s' is outside the implemented set for Whisper" % configs.get("decoder"))
        dec = dict(configs.get("decoder_conf", {}))
        need(dec, "input_layer", ("embed_learnable_pe",), "embed", "decoder_conf")
        need(dec, "activation_type", ("gelu",), "relu", "decoder_conf")
        need(dec, "normalize_before", (True,), True, "decoder_conf")
        need(dec, "src_attention", (True,), True, "decoder_conf")
        self.bidirectional = False
        self.dec_layers = int(dec.get("num_blocks", 6))
        self.rdec_layers = 0
        self.dec_heads = int(dec.get("attention_heads", 4))
        self.dec_ffn_dim = int(dec.get("linear_units", 2048))
        self.dec_ln_eps = float(dec.get("norm_eps", 1e-5))
        self.dec_flavor = 1
        self.dec_max_len = int(dec.get("max_len", 448))          # LearnablePositionalEncoding default (embedding.py:171)
        if self.d_model != self.dec_heads * 64:
            raise NotImplementedError("decoder head size must be 64")
        st = (configs.get("tokenizer_conf") or {}).get("special_tokens") or {}
        self.special_tokens = dict(st)
        self.sos = int(st.get("sot", self.vocab - 1))             # whisper.py:51-52
        self.eos = int(st.get("eot", self.vocab - 1))
        mc = dict(configs.get("model_conf", {}))
        self.reverse_weight = 0.0
        self.ctc_weight = float(mc.get("ctc_weight", 0.3))
        self.max_pos = int(enc.get("max_len", 1500))              # WhisperPositionalEncoding default (embedding.py:154)
        self.has_cmvn = False


def whisper_sinusoids(max_len: int, d: int) -> torch.Tensor:
    """WhisperPositionalEncoding table, wenet/models/transformer/embedding.py:150-164"""
    import math
    inc = math.log(10000) / (d // 2 - 1)
    inv = torch.exp(-inc * torch.arange(d // 2))
    st = torch.arange(max_len)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def _pack_whisper(spec: "ModelSpec", sd: Dict[str, torch.Tensor], precise: bool = False) -> Dict[str, torch.Tensor]:
    """Whisper (wenet/models/whisper/whisper.py; key names of TransformerEncoder / TransformerDecoder):
      * Conv1d(k=3) weights (out, in, 3) -> [out][(tap, in)] for the im2col GEMMs of whisper.cu
      * q/k/v fused; key_bias=False (attention.py:74-77) -> a zero bias slice
      * tie_word_embedding: output_layer.weight is the embedding matrix (decoder.py:283-311) - whatever the state_dict holds
        is packed, so tied and cloned checkpoints both work; a missing output bias packs as zeros"""
    d = spec.d_model
    out: Dict[str, torch.Tensor] = {}

    def f32(t):
        return t.detach().float().contiguous().cpu()

    def bf(t):
        return t.detach().float().to(torch.bfloat16).contiguous().cpu()

    def ebf(t, block=None):
        """encoder-side GEMM weight [N, K]: bf16, or (precise) [hi | hi | lo] per `block` columns of K (bf16x3)"""
        t = t.detach().float()
        if not precise:
            return bf(t)
        N, K = t.shape
        block = K if block is None else block
        return split3_weight(t.reshape(N * (K // block), block)).reshape(N, 3 * K).contiguous().cpu()

    def lin(dst, src, enc=False):
        w = sd[src + ".weight"]
        out[dst + ".w"] = ebf(w) if enc else bf(w)
        b = sd.get(src + ".bias")
        out[dst + ".b"] = f32(b) if b is not None else torch.zeros(w.shape[0])

    def norm(dst, src):
        out[dst + ".g"] = f32(sd[src + ".weight"])
        out[dst + ".b"] = f32(sd[src + ".bias"])

    def qkv(dst, a, enc=False):
        ws, bs = [], []
        for n in ("linear_q", "linear_k", "linear_v"):
            w = sd[a + "." + n + ".weight"]
            b = sd.get(a + "." + n + ".bias")
            ws.append(w)
            bs.append(b.float() if b is not None else torch.zeros(w.shape[0]))
        out[dst + ".w"] = ebf(torch.cat(ws, 0)) if enc else bf(torch.cat(ws, 0))
        out[dst + ".b"] = f32(torch.cat(bs, 0))

    w1 = sd["encoder.embed.conv.0.weight"]            # (d, idim, 3)
    out["wenc.conv1.w"] = ebf(w1.permute(0, 2, 1).reshape(d, 3 * spec.input_dim), block=spec.input_dim)
    out["wenc.conv1.b"] = f32(sd["encoder.embed.conv.0.bias"])
    w2 = sd["encoder.embed.conv.2.weight"]            # (d, d, 3)
    out["wenc.conv2.w"] = ebf(w2.permute(0, 2, 1).reshape(d, 3 * d), block=d)
    out["wenc.conv2.b"] = f32(sd["encoder.embed.conv.2.bias"])
    pe = sd.get("encoder.embed.pos_enc.pe")
    pe = whisper_sinusoids(spec.max_pos, d) if pe is None else pe.reshape(-1, d)
    assert pe.shape[0] == spec.max_pos, (pe.shape, spec.max_pos)
    out["wenc.pe"] = f32(pe)
    for i in range(spec.enc_layers):
        s_, t = "encoder.encoders.%d" % i, "wenc.%d" % i
        norm(t + ".norm1", s_ + ".norm1")
        norm(t + ".norm2", s_ + ".norm2")
        qkv(t + ".att.qkv", s_ + ".self_attn", enc=True)
        lin(t + ".att.out", s_ + ".self_attn.linear_out", enc=True)
        lin(t + ".ff.w1", s_ + ".feed_forward.w_1", enc=True)
        lin(t + ".ff.w2", s_ + ".feed_forward.w_2", enc=True)
    norm("after_norm", "encoder.after_norm")
    if spec.vocab > 0 and "ctc.ctc_lo.weight" in sd:
        lin("ctc", "ctc.ctc_lo", enc=True)
    if any(k.startswith("decoder.") for k in sd):
        dst, src = "dec.left", "decoder"
        out[dst + ".emb"] = f32(sd[src + ".embed.0.weight"])
        out[dst + ".pe"] = f32(sd[src + ".embed.1.pe"].reshape(-1, d))
        assert out[dst + ".pe"].shape[0] == spec.dec_max_len
        for i in range(spec.dec_layers):
            s_, t = "%s.decoders.%d" % (src, i), "%s.%d" % (dst, i)
            for n in ("norm1", "norm2", "norm3"):
                norm(t + "." + n, s_ + "." + n)
            qkv(t + ".sa.qkv", s_ + ".self_attn")
            lin(t + ".sa.out", s_ + ".self_attn.linear_out")
            a = s_ + ".src_attn"
            lin(t + ".ca.q", a + ".linear_q")
            wk, wv = sd[a + ".linear_k.weight"], sd[a + ".linear_v.weight"]
            bk, bv = sd.get(a + ".linear_k.bias"), sd.get(a + ".linear_v.bias")
            out[t + ".ca.kv.w"] = bf(torch.cat([wk, wv], 0))
            out[t + ".ca.kv.b"] = f32(torch.cat([bk.float() if bk is not None else torch.zeros(d),
                                                 bv.float() if bv is not None else torch.zeros(d)], 0))
            lin(t + ".ca.out", a + ".linear_out")
            lin(t + ".ff.w1", s_ + ".feed_forward.w_1")
            lin(t + ".ff.w2", s_ + ".feed_forward.w_2")
        norm(dst + ".after_norm", src + ".after_norm")
        lin(dst + ".out", src + ".output_layer")
    return out


def pack_state_dict(spec: ModelSpec, sd: Dict[str, torch.Tensor], precise: bool = False) -> Dict[str, torch.Tensor]:
    """Returns {lib tensor name: CPU tensor (fp32 or bf16)}.

    precise=True (wb_model_config.precise, the <= 1e-3 parity mode): every encoder / CTC GEMM weight is delivered as
    bf16 [N, 3K] = per K-block [hi | hi | lo] (bf16x3 against activations written as [hi | lo | hi]); the K-blocks are
    the d-wide channel groups the activations are split by (one block for a Linear, one per (kh, kw) tap for conv2,
    one per frequency bin for the embed Linear)."""
    if spec.arch == 1:
        return _pack_whisper(spec, sd, precise=precise)
    d, F1 = spec.d_model, (spec.input_dim - 3) // 2 + 1
    F2 = (F1 - 3) // 2 + 1
    out: Dict[str, torch.Tensor] = {}

    def f32(t):
        return t.detach().float().contiguous().cpu()

    def bf(t):
        return t.detach().float().to(torch.bfloat16).contiguous().cpu()

    def ebf(t, block=None):
        """encoder-side GEMM weight [N, K]: bf16, or (precise) [hi | hi | lo] per `block` columns of K"""
        t = t.detach().float()
        if not precise:
            return bf(t)
        N, K = t.shape
        block = K if block is None else block
        return split3_weight(t.reshape(N * (K // block), block)).reshape(N, 3 * K).contiguous().cpu()

    def lin(dst, src, enc=False):
        out[dst + ".w"] = ebf(sd[src + ".weight"]) if enc else bf(sd[src + ".weight"])
        out[dst + ".b"] = f32(sd[src + ".bias"])

    def norm(dst, src):
        out[dst + ".g"] = f32(sd[src + ".weight"])
        out[dst + ".b"] = f32(sd[src + ".bias"])

    if spec.has_cmvn:
        out["cmvn.mean"] = f32(sd["encoder.global_cmvn.mean"])
        out["cmvn.istd"] = f32(sd["encoder.global_cmvn.istd"])
    w1 = sd["encoder.embed.conv.0.weight"]          # (d, 1, 3, 3)
    out["embed.conv1.w"] = f32(w1.reshape(d, 9).t())
    out["embed.conv1.b"] = f32(sd["encoder.embed.conv.0.bias"])
    w2 = sd["encoder.embed.conv.2.weight"]          # (d, d, 3, 3) -> (d, kh, kw, c_in)
    out["embed.conv2.w"] = ebf(w2.permute(0, 2, 3, 1).reshape(d, 9 * d), block=d)
    out["embed.conv2.b"] = f32(sd["encoder.embed.conv.2.bias"])
    wo = sd["encoder.embed.out.0.weight"]           # (d, c * F2 + f) -> (d, f * d + c)
    assert wo.shape[1] == d * F2, "embed.out expects %d input features, got %d" % (d * F2, wo.shape[1])
    out["embed.out.w"] = ebf(wo.view(d, d, F2).permute(0, 2, 1).reshape(d, F2 * d), block=d)
    out["embed.out.b"] = f32(sd["encoder.embed.out.0.bias"])
    pe = sd.get("encoder.embed.pos_enc.pe")
    pe = sinusoid_pe(spec.max_pos, d) if pe is None else pe.reshape(-1, d)[:spec.max_pos]
    out["embed.pe"] = f32(pe)
    for i in range(spec.enc_layers):
        s, t = "encoder.encoders.%d" % i, "enc.%d" % i
        for n in ("norm_ff_macaron", "norm_mha", "norm_conv", "norm_ff", "norm_final"):
            norm(t + "." + n, s + "." + n)
        lin(t + ".ffm.w1", s + ".feed_forward_macaron.w_1", enc=True)
        lin(t + ".ffm.w2", s + ".feed_forward_macaron.w_2", enc=True)
        lin(t + ".ff.w1", s + ".feed_forward.w_1", enc=True)
        lin(t + ".ff.w2", s + ".feed_forward.w_2", enc=True)
        a = s + ".self_attn"
        out[t + ".att.qkv.w"] = ebf(torch.cat([sd[a + ".linear_q.weight"], sd[a + ".linear_k.weight"],
                                              sd[a + ".linear_v.weight"]], 0))
        out[t + ".att.qkv.b"] = f32(torch.cat([sd[a + ".linear_q.bias"], sd[a + ".linear_k.bias"],
                                               sd[a + ".linear_v.bias"]], 0))
        lin(t + ".att.out", a + ".linear_out", enc=True)
        out[t + ".att.pos.w3"] = split3_weight(sd[a + ".linear_pos.weight"]).cpu()
        out[t + ".att.pos_u"] = f32(sd[a + ".pos_bias_u"].reshape(-1))
        out[t + ".att.pos_v"] = f32(sd[a + ".pos_bias_v"].reshape(-1))
        c = s + ".conv_module"
        pw1_w = sd[c + ".pointwise_conv1.weight"].reshape(2 * d, d)
        pw1_b = sd[c + ".pointwise_conv1.bias"]
        wi, bi = interleave_glu(pw1_w, pw1_b)
        out[t + ".conv.pw1.w"], out[t + ".conv.pw1.b"] = ebf(wi), f32(bi)
        out[t + ".conv.pad_vec"] = f32(pw1_b[:d].float() * torch.sigmoid(pw1_b[d:].float()))
        out[t + ".conv.dw.w"] = f32(sd[c + ".depthwise_conv.weight"].reshape(d, spec.cnn_kernel))
        out[t + ".conv.dw.b"] = f32(sd[c + ".depthwise_conv.bias"])
        if spec.cnn_norm == "layer_norm":
            norm(t + ".conv.norm", c + ".norm")
        else:
            scale = sd[c + ".norm.weight"].float() / torch.sqrt(sd[c + ".norm.running_var"].float() + spec.ln_eps)
            out[t + ".conv.norm.g"] = f32(scale)
            out[t + ".conv.norm.b"] = f32(sd[c + ".norm.bias"].float() - sd[c + ".norm.running_mean"].float() * scale)
        out[t + ".conv.pw2.w"] = ebf(sd[c + ".pointwise_conv2.weight"].reshape(d, d))
        out[t + ".conv.pw2.b"] = f32(sd[c + ".pointwise_conv2.bias"])
    norm("after_norm", "encoder.after_norm")
    if spec.vocab > 0:      # encoder-only handles (plugin: a stand-alone ConformerEncoder) carry no CTC head
        lin("ctc", "ctc.ctc_lo", enc=True)

    def decoder(dst, src, n_layers):
        out[dst + ".emb"] = f32(sd[src + ".embed.0.weight"])
        for i in range(n_layers):
            s, t = "%s.decoders.%d" % (src, i), "%s.%d" % (dst, i)
            for n in ("norm1", "norm2", "norm3"):
                norm(t + "." + n, s + "." + n)
            a = s + ".self_attn"
            out[t + ".sa.qkv.w"] = bf(torch.cat([sd[a + ".linear_q.weight"], sd[a + ".linear_k.weight"],
                                                 sd[a + ".linear_v.weight"]], 0))
            out[t + ".sa.qkv.b"] = f32(torch.cat([sd[a + ".linear_q.bias"], sd[a + ".linear_k.bias"],
                                                  sd[a + ".linear_v.bias"]], 0))
            lin(t + ".sa.out", a + ".linear_out")
            a = s + ".src_attn"
            lin(t + ".ca.q", a + ".linear_q")
            out[t + ".ca.kv.w"] = bf(torch.cat([sd[a + ".linear_k.weight"], sd[a + ".linear_v.weight"]], 0))
            out[t + ".ca.kv.b"] = f32(torch.cat([sd[a + ".linear_k.bias"], sd[a + ".linear_v.bias"]], 0))
            lin(t + ".ca.out", a + ".linear_out")
            lin(t + ".ff.w1", s + ".feed_forward.w_1")
            lin(t + ".ff.w2", s + ".feed_forward.w_2")
        norm(dst + ".after_norm", src + ".after_norm")
        lin(dst + ".out", src + ".output_layer")

    has_dec = any(k.startswith("decoder.") for k in sd)
    if has_dec:
        if spec.bidirectional:
            decoder("dec.left", "decoder.left_decoder", spec.dec_layers)
            if spec.rdec_layers > 0:
                decoder("dec.right", "decoder.right_decoder", spec.rdec_layers)
        else:
            decoder("dec.left", "decoder", spec.dec_layers)
    return out


class DeviceModel:
    """Owns a wb_model handle (weights resident in HBM)."""

    def __init__(self, spec: ModelSpec, sd: Dict[str, torch.Tensor], with_decoder: bool = True, precise: bool = False):
        self.spec = spec
        self.precise = bool(precise)
        lib = _lib.load()
        has_dec = with_decoder and any(k.startswith("decoder.") for k in sd)
        cfg = WbModelConfig(
            input_dim=spec.input_dim, d_model=spec.d_model, heads=spec.heads, ffn_dim=spec.ffn_dim,
            enc_layers=spec.enc_layers, cnn_kernel=spec.cnn_kernel, cnn_causal=int(spec.cnn_causal),
            cnn_norm=0 if spec.cnn_norm == "layer_norm" else 1, vocab=spec.vocab,
            dec_layers=spec.dec_layers if has_dec else 0, rdec_layers=spec.rdec_layers if has_dec else 0,
            dec_heads=spec.dec_heads, dec_ffn_dim=spec.dec_ffn_dim, max_pos=spec.max_pos,
            has_cmvn=int(spec.has_cmvn), precise=int(self.precise), ln_eps=spec.ln_eps, dec_ln_eps=spec.dec_ln_eps,
            arch=int(spec.arch), dec_flavor=int(spec.dec_flavor), dec_max_len=int(spec.dec_max_len))
        self._h = C.c_void_p()
        check(lib.wb_model_create(C.byref(self._h), C.byref(cfg)), "wb_model_create")
        packed = pack_state_dict(spec, sd, precise=self.precise)
        for name, t in packed.items():
            if not has_dec and name.startswith("dec."):
                continue
            if t.dtype == torch.bfloat16:
                arr = t.view(torch.int16).numpy()
                dt = WB_BF16
            else:
                arr = t.numpy()
                dt = WB_F32
            arr = np.ascontiguousarray(arr)
            check(lib.wb_model_set_tensor(self._h, name.encode(), C.c_void_p(arr.ctypes.data), dt, arr.size),
                  "wb_model_set_tensor(%s)" % name)
        check(lib.wb_model_finalize(self._h, cur_stream()), "wb_model_finalize")
        self.has_decoder = has_dec

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if self._h:
                _lib.load().wb_model_destroy(self._h)
        except Exception:
            pass
