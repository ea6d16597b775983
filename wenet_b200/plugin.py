"""Drop-in integration with an importable WeNet (the reference): `install()` rebinds the reference's own plug-in
points so `wenet/bin/recognize.py`, `wenet.load_model()` and any code holding a reference-built model run the B200
path unmodified:

  * wenet.utils.init_model.WENET_ENCODER_CLASSES["conformer"] (init_model.py:52-64) -> a subclass of the reference
    ConformerEncoder whose eval-mode `forward`, `forward_chunk` and `forward_chunk_by_chunk` (encoder.py:122-362)
    call libwenet_b200.so.  Hence `model.encoder(...)`, `model.encoder.forward_chunk(...)`,
    `model.forward_encoder_chunk(...)` (asr_model.py:385-426) and `model._forward_encoder(...)` are B200 calls.
  * WENET_CTC_CLASSES["ctc"] (init_model.py:72-74) -> a CTC subclass whose `log_softmax` (ctc.py:73-81), and with it
    `model.ctc_activation` (asr_model.py:428-438), is the tcgen05 GEMM + log-softmax kernel.
  * WENET_MODEL_CLASSES["asr_model"] (init_model.py:88-97) -> a subclass of the reference ASRModel whose `decode`,
    `ctc_logprobs` and `forward_attention_decoder` call the library (`transcribe` reaches `decode`).
    All three keep the reference module tree, so `load_checkpoint` / `load_state_dict` / `state_dict` / `.to()` see the
    reference's parameters; the packed device weights are (re)built lazily from `state_dict()` when parameters change.
  * wenet.dataset.processor.compute_fbank (processor.py:226-256; looked up by attribute at dataset.py:96 and
    cli/model.py:58) -> the fused CUDA fbank (dither == 0 in the main process; see `_fbank_dropin`).
  * WENET_MODEL_CLASSES["whisper"] (init_model.py:91) -> a subclass of the reference Whisper whose `decode` runs the B200
    Whisper path (wenet_b200/whisper.py), and processor.compute_log_mel_spectrogram (processor.py:320-369) -> the CUDA
    log-mel kernel.

Training mode (`module.training`) always takes the reference implementation (autograd); eval mode needs CUDA tensors and
raises otherwise - there is no CPU fallback.  Configurations outside the implemented set raise NotImplementedError when
the encoder is constructed.  This module is the only one that imports `wenet`.
"""
import os
import weakref
from typing import Dict

import torch

from . import _lib
from ._lib import check, cur_stream, ptr
from .asr_model import B200ASRModel
from .fbank import compute_fbank as b200_compute_fbank


def encoder_configs_from_module(enc) -> dict:
    """train.yaml `encoder_conf` subset (+ input_dim / cmvn) recovered from a constructed reference ConformerEncoder;
    raises NotImplementedError for anything outside the implemented set."""
    name = type(enc).__mro__
    if not any(c.__name__ == "ConformerEncoder" for c in name):
        raise NotImplementedError("encoder class %s is outside the implemented set (ConformerEncoder)" % type(enc).__name__)
    if len(enc.encoders) == 0:
        raise NotImplementedError("encoder without layers")
    layer = enc.encoders[0]
    emb = type(enc.embed).__name__
    if emb != "Conv2dSubsampling4":
        raise NotImplementedError("input layer %s is outside the implemented set (conv2d)" % emb)
    if type(enc.embed.pos_enc).__name__ != "RelPositionalEncoding":
        raise NotImplementedError("positional encoding %s is outside the implemented set (rel_pos)"
                                  % type(enc.embed.pos_enc).__name__)
    att = type(layer.self_attn).__name__
    if att != "RelPositionMultiHeadedAttention":
        raise NotImplementedError("attention %s is outside the implemented set (rel_selfattn)" % att)
    if layer.conv_module is None or layer.feed_forward_macaron is None:
        raise NotImplementedError("conformer layer without conv module / macaron FFN")
    if type(layer.feed_forward).__name__ != "PositionwiseFeedForward":
        raise NotImplementedError("mlp type %s is outside the implemented set" % type(layer.feed_forward).__name__)
    act = type(layer.feed_forward.activation).__name__
    if act != "SiLU":
        raise NotImplementedError("activation %s is outside the implemented set (swish)" % act)
    if type(enc.after_norm).__name__ != "LayerNorm" or type(layer.norm_ff).__name__ != "LayerNorm":
        raise NotImplementedError("layer_norm_type %s is outside the implemented set (layer_norm)"
                                  % type(enc.after_norm).__name__)
    if not enc.normalize_before:
        raise NotImplementedError("post-norm encoders are outside the implemented set")
    sa = layer.self_attn
    if getattr(sa, "h_kv", sa.h) != sa.h:
        raise NotImplementedError("grouped-query attention (n_kv_head) is outside the implemented set")
    conv = layer.conv_module
    d = enc.output_size()
    # input_dim: GlobalCMVN knows it; otherwise only F2 = in_features / d is visible and input dims 4*F2+3 .. 4*F2+6 all
    # map to it - assume the smallest even one (80 for F2 = 19); a batch of another width fails loudly in encode()
    if enc.global_cmvn is not None:
        input_dim = int(enc.global_cmvn.mean.numel())
    else:
        input_dim = 4 * (enc.embed.out[0].in_features // d) + 4
    return {
        "input_dim": input_dim,
        "cmvn": "global_cmvn" if enc.global_cmvn is not None else None,
        "encoder": "conformer",
        "encoder_conf": dict(
            output_size=d, attention_heads=int(sa.h),
            linear_units=int(layer.feed_forward.w_1.out_features), num_blocks=len(enc.encoders),
            input_layer="conv2d", pos_enc_layer_type="rel_pos", selfattention_layer_type="rel_selfattn",
            activation_type="swish", normalize_before=True, use_cnn_module=True,
            cnn_module_kernel=int(conv.depthwise_conv.kernel_size[0]), causal=bool(conv.lorder > 0),
            cnn_module_norm="layer_norm" if conv.use_layer_norm else "batch_norm",
            use_dynamic_chunk=bool(enc.use_dynamic_chunk), static_chunk_size=int(enc.static_chunk_size),
            norm_eps=float(enc.after_norm.eps)),
    }


def configs_from_reference_model(model) -> dict:
    """Reconstruct the train.yaml subset this build needs from a constructed reference ASRModel."""
    cfg = encoder_configs_from_module(model.encoder)
    cfg["output_dim"] = int(model.vocab_size)
    dec = model.decoder
    dname = type(dec).__name__
    if dname == "BiTransformerDecoder":
        left = dec.left_decoder
        cfg["decoder"] = "bitransformer"
        cfg["decoder_conf"] = dict(attention_heads=int(left.decoders[0].self_attn.h),
                                   linear_units=int(left.decoders[0].feed_forward.w_1.out_features),
                                   num_blocks=len(left.decoders), r_num_blocks=len(dec.right_decoder.decoders),
                                   norm_eps=float(left.after_norm.eps))
    elif dname == "TransformerDecoder":
        cfg["decoder"] = "transformer"
        cfg["decoder_conf"] = dict(attention_heads=int(dec.decoders[0].self_attn.h),
                                   linear_units=int(dec.decoders[0].feed_forward.w_1.out_features),
                                   num_blocks=len(dec.decoders), norm_eps=float(dec.after_norm.eps))
    else:
        raise NotImplementedError("decoder class %s is outside the implemented set" % dname)
    cfg["model_conf"] = dict(ctc_weight=float(model.ctc_weight), reverse_weight=float(model.reverse_weight))
    return cfg


def configs_from_reference_whisper(model) -> dict:
    """train.yaml subset of a constructed reference Whisper (wenet/models/whisper/whisper.py); raises NotImplementedError for
    anything outside the implemented set (conv1d2 / abs_pos_whisper / gelu / pre-norm, embed_learnable_pe decoder)."""
    enc, dec = model.encoder, model.decoder
    if type(enc).__name__ != "TransformerEncoder" or type(enc.embed).__name__ != "Conv1dSubsampling2":
        raise NotImplementedError("Whisper encoder %s / %s is outside the implemented set (transformer / conv1d2)"
                                  % (type(enc).__name__, type(enc.embed).__name__))
    if type(enc.embed.pos_enc).__name__ != "WhisperPositionalEncoding":
        raise NotImplementedError("positional encoding %s is outside the implemented set (abs_pos_whisper)"
                                  % type(enc.embed.pos_enc).__name__)
    if type(dec).__name__ != "TransformerDecoder" or type(dec.embed[1]).__name__ != "LearnablePositionalEncoding":
        raise NotImplementedError("Whisper decoder outside the implemented set (transformer / embed_learnable_pe)")
    el, dl = enc.encoders[0], dec.decoders[0]
    for lyr in (el, dl):
        if type(lyr.feed_forward.activation).__name__ != "GELU" or not lyr.normalize_before:
            raise NotImplementedError("Whisper layers must be pre-norm with GELU")
    d = enc.output_size()
    return {
        "input_dim": int(enc.embed.conv[0].in_channels), "output_dim": int(model.vocab_size), "cmvn": None,
        "encoder": "transformer",
        "encoder_conf": dict(output_size=d, attention_heads=int(el.self_attn.h), linear_units=int(el.feed_forward.w_1.out_features),
                             num_blocks=len(enc.encoders), input_layer="conv1d2", pos_enc_layer_type="abs_pos_whisper",
                             activation_type="gelu", normalize_before=True, key_bias=el.self_attn.linear_k.bias is not None,
                             norm_eps=float(enc.after_norm.eps), max_len=int(enc.embed.pos_enc.pe.shape[1]),
                             use_dynamic_chunk=bool(enc.use_dynamic_chunk), static_chunk_size=max(int(enc.static_chunk_size), 0)),
        "decoder": "transformer",
        "decoder_conf": dict(attention_heads=int(dl.self_attn.h), linear_units=int(dl.feed_forward.w_1.out_features),
                             num_blocks=len(dec.decoders), input_layer="embed_learnable_pe", activation_type="gelu",
                             normalize_before=True, src_attention=True, norm_eps=float(dec.after_norm.eps),
                             max_len=int(dec.embed[1].pe.shape[1])),
        "tokenizer": "whisper", "tokenizer_conf": dict(special_tokens=dict(model.special_tokens)),
        "model": "whisper", "model_conf": dict(ctc_weight=float(model.ctc_weight)),
    }


def wrap(model, device=None, precise: bool = False) -> B200ASRModel:
    """B200 core object sharing the weights of a loaded reference model."""
    core = B200ASRModel.from_reference(model, configs_from_reference_model(model), device=device, precise=precise)
    core.sos, core.eos = model.sos_symbol(), model.eos_symbol()
    return core


def _version(module) -> int:
    return sum(int(p._version) for p in module.parameters()) + sum(int(b._version) for b in module.buffers())


def _cuda_device_of(module):
    dev = next(module.parameters()).device
    if dev.type != "cuda":
        raise _lib.WbError("the B200 path needs the module on a CUDA device (no CPU fallback); call .to('cuda') first")
    return dev


_classes = None
_orig_compute_fbank = None
_orig_compute_logmel = None


def _logmel_dropin(sample, n_fft=400, hop_length=160, num_mel_bins=80, padding=0, pad_or_trim: bool = False,
                   max_duration: int = 30):
    """wenet.dataset.processor.compute_log_mel_spectrogram after install() (processor.py:320-369; looked up by attribute at
    dataset.py and cli/model.py): the CUDA log-mel kernel in the main process, the reference's own function inside DataLoader
    workers (CUDA cannot be initialised in a forked worker)."""
    if torch.utils.data.get_worker_info() is not None:
        return _orig_compute_logmel(sample, n_fft=n_fft, hop_length=hop_length, num_mel_bins=num_mel_bins, padding=padding,
                                    pad_or_trim=pad_or_trim, max_duration=max_duration)
    from .whisper import compute_log_mel_spectrogram as b200_logmel
    return b200_logmel(sample, n_fft=n_fft, hop_length=hop_length, num_mel_bins=num_mel_bins, padding=padding,
                       pad_or_trim=pad_or_trim, max_duration=max_duration)


def _fbank_dropin(sample, num_mel_bins=23, frame_length=25, frame_shift=10, dither=0.0, window_type="povey"):
    """wenet.dataset.processor.compute_fbank after install().  The CUDA kernel serves decoding in the main process
    (dither == 0, recognize.py:225-226).  Two cases keep the reference's own function, because the CUDA path cannot
    serve them at all: dither != 0 (a training-time augmentation drawing torch RNG noise) and DataLoader worker
    processes (CUDA cannot be initialised in a forked worker)."""
    in_worker = torch.utils.data.get_worker_info() is not None
    if dither != 0.0 or in_worker:
        return _orig_compute_fbank(sample, num_mel_bins=num_mel_bins, frame_length=frame_length,
                                   frame_shift=frame_shift, dither=dither, window_type=window_type)
    return b200_compute_fbank(sample, num_mel_bins=num_mel_bins, frame_length=frame_length, frame_shift=frame_shift,
                              dither=dither, window_type=window_type)


def _make_classes():
    from wenet.models.transformer.asr_model import ASRModel
    from wenet.models.transformer.ctc import CTC
    from wenet.models.transformer.encoder import ConformerEncoder
    from wenet.models.transformer.search import DecodeResult as RefDecodeResult

    class B200ConformerEncoderPlugin(ConformerEncoder):
        """Reference ConformerEncoder (same constructor, parameters and state_dict keys); eval-mode inference runs on
        libwenet_b200.so.  Inside a B200ASRModelPlugin it shares the model's packed weights, on its own it packs an
        encoder-only copy."""

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            encoder_configs_from_module(self)        # unsupported configurations fail here, at construction

        def _b200(self) -> B200ASRModel:
            owner = self.__dict__.get("_b200_owner")
            owner = owner() if owner is not None else None
            if owner is not None and getattr(owner, "encoder", None) is self:
                return owner._b200()
            dev = _cuda_device_of(self)
            ver = _version(self)
            core = self.__dict__.get("_b200_core")
            if core is None or self.__dict__.get("_b200_ver") != ver or core.device != dev:
                cfg = dict(encoder_configs_from_module(self), output_dim=0, decoder="transformer", decoder_conf={})
                sd = {"encoder." + k: v.detach().cpu() for k, v in self.state_dict().items()}
                core = B200ASRModel(cfg, sd, device=dev, with_decoder=False)
                self.__dict__["_b200_core"], self.__dict__["_b200_ver"] = core, ver
            return core

        def forward(self, xs, xs_lens, decoding_chunk_size: int = 0, num_decoding_left_chunks: int = -1):
            if self.training:
                return super().forward(xs, xs_lens, decoding_chunk_size, num_decoding_left_chunks)
            return self._b200().encoder.forward(xs, xs_lens, decoding_chunk_size, num_decoding_left_chunks)

        def forward_chunk(self, xs, offset, required_cache_size, att_cache=torch.zeros(0, 0, 0, 0),
                          cnn_cache=torch.zeros(0, 0, 0, 0), att_mask=torch.ones((0, 0, 0), dtype=torch.bool)):
            if self.training:
                return super().forward_chunk(xs, offset, required_cache_size, att_cache, cnn_cache, att_mask)
            return self._b200().encoder.forward_chunk(xs, offset, required_cache_size, att_cache, cnn_cache)

        def forward_chunk_by_chunk(self, xs, decoding_chunk_size: int, num_decoding_left_chunks: int = -1):
            if self.training:
                return super().forward_chunk_by_chunk(xs, decoding_chunk_size, num_decoding_left_chunks)
            return self._b200().encoder.forward_chunk_by_chunk(xs, decoding_chunk_size, num_decoding_left_chunks)

    class B200CTCPlugin(CTC):
        """Reference CTC module; eval-mode log_softmax = tcgen05 GEMM (fp32 logits) + the log-softmax kernel."""

        def log_softmax(self, hs_pad: torch.Tensor) -> torch.Tensor:
            if self.training:
                return super().log_softmax(hs_pad)
            if not hs_pad.is_cuda:
                raise _lib.WbError("hs_pad must be a CUDA tensor (no CPU fallback)")
            _cuda_device_of(self)
            lib = _lib.load()
            ver = _version(self)
            if self.__dict__.get("_b200_ver") != ver or self.__dict__["_b200_w"].device != hs_pad.device:
                self.__dict__["_b200_w"] = self.ctc_lo.weight.detach().to(hs_pad.device, torch.bfloat16).contiguous()
                self.__dict__["_b200_b"] = self.ctc_lo.bias.detach().to(hs_pad.device, torch.float32).contiguous()
                self.__dict__["_b200_ver"] = ver
            w, b = self.__dict__["_b200_w"], self.__dict__["_b200_b"]
            V, d = w.shape
            lead = hs_pad.shape[:-1]
            x = hs_pad.reshape(-1, d).to(torch.float32).contiguous()
            R = x.shape[0]
            ldl = (V + 7) // 8 * 8
            out = torch.empty(max(R, 1), ldl, device=x.device, dtype=torch.float32)
            if R > 0:
                with torch.cuda.device(x.device):
                    a = torch.empty(R, d, device=x.device, dtype=torch.bfloat16)
                    tv = torch.empty(R, 1, device=x.device, dtype=torch.float32)
                    ti = torch.empty(R, 1, device=x.device, dtype=torch.int32)
                    st = cur_stream()
                    check(lib.wb_op_cast_bf16(ptr(x), d, R, d, ptr(a), d, 0, st), "wb_op_cast_bf16")
                    check(lib.wb_op_gemm(ptr(a), d, ptr(w), R, V, d, ptr(b), 5, 1.0, ptr(out), ldl, 0, st), "wb_op_gemm")
                    check(lib.wb_op_logsoftmax_topk(ptr(out), ldl, R, V, 0, 0.0, 1, ptr(tv), ptr(ti), st),
                          "wb_op_logsoftmax_topk")
            return out[:R, :V].reshape(*lead, V)

    class B200ASRModelPlugin(ASRModel):
        """Reference ASRModel (same constructor, same parameters / state_dict) whose inference runs on
        libwenet_b200.so.  Training methods are inherited unchanged."""

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            if isinstance(self.encoder, B200ConformerEncoderPlugin):
                self.encoder.__dict__["_b200_owner"] = weakref.ref(self)     # share one packed weight set
            else:
                encoder_configs_from_module(self.encoder)                    # raises: unsupported encoder class

        def _b200(self) -> B200ASRModel:
            dev = _cuda_device_of(self)
            core = self.__dict__.get("_b200_core")
            ver = _version(self)
            if core is None or self.__dict__.get("_b200_ver") != ver or core.device != dev:
                # WENET_B200_PRECISE=1 (or model.b200_precise = True) selects the <= 1e-3 parity mode
                precise = self.__dict__.get("b200_precise", os.environ.get("WENET_B200_PRECISE", "0") == "1")
                core = wrap(self, device=dev, precise=bool(precise))
                self.__dict__["_b200_core"], self.__dict__["_b200_ver"] = core, ver
            return core

        def ctc_logprobs(self, encoder_out, blank_penalty: float = 0.0, blank_id: int = 0):
            return self._b200().ctc_logprobs(encoder_out, blank_penalty, blank_id)

        def forward_attention_decoder(self, hyps, hyps_lens, encoder_out, reverse_weight: float = 0):
            return self._b200().forward_attention_decoder(hyps, hyps_lens, encoder_out, reverse_weight)

        def decode(self, methods, speech, speech_lengths, beam_size=1, decoding_chunk_size=-1,
                   num_decoding_left_chunks=-1, ctc_weight=0.0, simulate_streaming=False, reverse_weight=0.0,
                   context_graph=None, blank_id=0, blank_penalty=0.0, length_penalty=0.0, infos=None):
            res = self._b200().decode(methods, speech, speech_lengths, beam_size, decoding_chunk_size,
                                      num_decoding_left_chunks, ctc_weight, simulate_streaming, reverse_weight,
                                      context_graph, blank_id, blank_penalty, length_penalty, infos)
            return {k: [RefDecodeResult(tokens=r.tokens, score=r.score, confidence=r.confidence,
                                        tokens_confidence=r.tokens_confidence, times=r.times, nbest=r.nbest,
                                        nbest_scores=r.nbest_scores, nbest_times=r.nbest_times) for r in v]
                    for k, v in res.items()}

    from wenet.models.whisper.whisper import Whisper
    from wenet.utils.common import WHISPER_LANGS as REF_WHISPER_LANGS

    class B200WhisperPlugin(Whisper):
        """Reference Whisper (same constructor / parameters / state_dict); `decode` (and with it `transcribe`, the
        wenet.cli path) runs log-mel features -> TransformerEncoder -> attention_beam_search on libwenet_b200.so."""

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            configs_from_reference_whisper(self)      # unsupported configurations fail here, at construction

        def _b200(self):
            from .whisper import B200Whisper
            dev = _cuda_device_of(self)
            core = self.__dict__.get("_b200_core")
            ver = _version(self)
            if core is None or self.__dict__.get("_b200_ver") != ver or core.device != dev:
                sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
                precise = self.__dict__.get("b200_precise", os.environ.get("WENET_B200_PRECISE", "0") == "1")
                core = B200Whisper(configs_from_reference_whisper(self), sd, device=dev, lang_table=REF_WHISPER_LANGS,
                                   precise=bool(precise))
                self.__dict__["_b200_core"], self.__dict__["_b200_ver"] = core, ver
            return core

        def decode(self, methods, speech, speech_lengths, beam_size=1, decoding_chunk_size=-1,
                   num_decoding_left_chunks=-1, ctc_weight=0.0, simulate_streaming=False, reverse_weight=0.0,
                   context_graph=None, blank_id=0, blank_penalty=0.0, length_penalty=0.0, infos=None):
            res = self._b200().decode(methods, speech, speech_lengths, beam_size, decoding_chunk_size,
                                      num_decoding_left_chunks, ctc_weight, simulate_streaming, reverse_weight,
                                      context_graph, blank_id, blank_penalty, length_penalty, infos)
            return {k: [RefDecodeResult(tokens=r.tokens, score=r.score, confidence=r.confidence,
                                        tokens_confidence=r.tokens_confidence, times=r.times, nbest=r.nbest,
                                        nbest_scores=r.nbest_scores, nbest_times=r.nbest_times) for r in v]
                    for k, v in res.items()}

    return B200ASRModelPlugin, B200ConformerEncoderPlugin, B200CTCPlugin, B200WhisperPlugin


def install(fbank: bool = True):
    """Rebind WeNet's registries (needs `wenet` importable).  Idempotent.  Returns the plugin model class.
    fbank=False leaves wenet.dataset.processor.compute_fbank alone."""
    global _classes, _orig_compute_fbank, _orig_compute_logmel
    import wenet.dataset.processor as processor
    from wenet.utils import init_model as im
    if _classes is None:
        _classes = _make_classes()
    model_cls, enc_cls, ctc_cls, whisper_cls = _classes
    im.WENET_MODEL_CLASSES["asr_model"] = model_cls
    im.WENET_MODEL_CLASSES["whisper"] = whisper_cls
    im.WENET_ENCODER_CLASSES["conformer"] = enc_cls
    im.WENET_CTC_CLASSES["ctc"] = ctc_cls
    if fbank and processor.compute_fbank is not _fbank_dropin:
        _orig_compute_fbank = processor.compute_fbank
        processor.compute_fbank = _fbank_dropin
    if fbank and processor.compute_log_mel_spectrogram is not _logmel_dropin:
        _orig_compute_logmel = processor.compute_log_mel_spectrogram
        processor.compute_log_mel_spectrogram = _logmel_dropin
    return model_cls


def uninstall():
    """Restore the reference's own classes / compute_fbank (used by the tests)."""
    global _orig_compute_fbank, _orig_compute_logmel
    import wenet.dataset.processor as processor
    from wenet.models.transformer.asr_model import ASRModel
    from wenet.models.transformer.ctc import CTC
    from wenet.models.transformer.encoder import ConformerEncoder
    from wenet.utils import init_model as im
    from wenet.models.whisper.whisper import Whisper
    im.WENET_MODEL_CLASSES["asr_model"] = ASRModel
    im.WENET_MODEL_CLASSES["whisper"] = Whisper
    im.WENET_ENCODER_CLASSES["conformer"] = ConformerEncoder
    im.WENET_CTC_CLASSES["ctc"] = CTC
    if _orig_compute_fbank is not None and processor.compute_fbank is _fbank_dropin:
        processor.compute_fbank = _orig_compute_fbank
        _orig_compute_fbank = None
    if _orig_compute_logmel is not None and processor.compute_log_mel_spectrogram is _logmel_dropin:
        processor.compute_log_mel_spectrogram = _orig_compute_logmel
        _orig_compute_logmel = None
