"""Drop-in integration with an installed/importable WeNet (the reference): `install()` rebinds
the reference's own plug-in points so `wenet/bin/recognize.py` and `wenet.load_model()` run the B200
path unmodified:

  * wenet.utils.init_model.WENET_MODEL_CLASSES["asr_model"]  (init_model.py:88-97) -> a subclass of
    the reference ASRModel whose inference methods (decode, _forward_encoder, ctc_logprobs,
    forward_attention_decoder, encoder.forward_chunk via the core) call libwenet_b200.so.  It keeps the
    reference module tree, so `load_checkpoint` / `load_state_dict` / `state_dict` see the reference key
    names; the packed device weights are (re)built lazily from `state_dict()`.
  * wenet.dataset.processor.compute_fbank (processor.py:226-256; looked up by name at
    dataset.py:96 and cli/model.py:58) -> the fused CUDA fbank.

This module is the only one that imports `wenet`; it is optional (the core API in asr_model.py has no
dependency on the reference).
"""
from typing import Dict

import torch

from . import _lib
from .asr_model import B200ASRModel
from .fbank import compute_fbank as b200_compute_fbank


def configs_from_reference_model(model) -> dict:
    """Reconstruct the train.yaml subset this build needs from a constructed reference ASRModel."""
    enc = model.encoder
    name = type(enc).__name__
    if name != "ConformerEncoder":
        raise NotImplementedError("encoder class %s is outside the implemented set (ConformerEncoder)" % name)
    layer = enc.encoders[0]
    emb = type(enc.embed).__name__
    if emb != "Conv2dSubsampling4":
        raise NotImplementedError("input layer %s is outside the implemented set (conv2d)" % emb)
    att = type(layer.self_attn).__name__
    if att != "RelPositionMultiHeadedAttention":
        raise NotImplementedError("attention %s is outside the implemented set (rel_selfattn)" % att)
    if layer.conv_module is None or layer.feed_forward_macaron is None:
        raise NotImplementedError("conformer layer without conv module / macaron FFN")
    act = type(layer.feed_forward.activation).__name__
    if act != "SiLU":
        raise NotImplementedError("activation %s is outside the implemented set (swish)" % act)
    conv = layer.conv_module
    d = enc.output_size()
    # input_dim from the embed Linear: in_features = d * (((idim - 1) // 2 - 1) // 2); global_cmvn knows it too
    if enc.global_cmvn is not None:
        input_dim = int(enc.global_cmvn.mean.numel())
    else:
        f2 = enc.embed.out[0].in_features // d
        input_dim = 4 * f2 + 4   # smallest idim giving F2 (80 -> 19)
    cfg = {
        "input_dim": input_dim,
        "output_dim": int(model.vocab_size),
        "cmvn": "global_cmvn" if enc.global_cmvn is not None else None,
        "encoder": "conformer",
        "encoder_conf": dict(
            output_size=d, attention_heads=int(layer.self_attn.h),
            linear_units=int(layer.feed_forward.w_1.out_features), num_blocks=len(enc.encoders),
            input_layer="conv2d", pos_enc_layer_type="rel_pos", selfattention_layer_type="rel_selfattn",
            activation_type="swish", normalize_before=bool(enc.normalize_before), use_cnn_module=True,
            cnn_module_kernel=int(conv.depthwise_conv.kernel_size[0]), causal=bool(conv.lorder > 0),
            cnn_module_norm="layer_norm" if conv.use_layer_norm else "batch_norm",
            use_dynamic_chunk=bool(enc.use_dynamic_chunk), static_chunk_size=int(enc.static_chunk_size)),
    }
    dec = model.decoder
    dname = type(dec).__name__
    if dname == "BiTransformerDecoder":
        left = dec.left_decoder
        cfg["decoder"] = "bitransformer"
        cfg["decoder_conf"] = dict(attention_heads=int(left.decoders[0].self_attn.h),
                                   linear_units=int(left.decoders[0].feed_forward.w_1.out_features),
                                   num_blocks=len(left.decoders), r_num_blocks=len(dec.right_decoder.decoders))
    elif dname == "TransformerDecoder":
        cfg["decoder"] = "transformer"
        cfg["decoder_conf"] = dict(attention_heads=int(dec.decoders[0].self_attn.h),
                                   linear_units=int(dec.decoders[0].feed_forward.w_1.out_features),
                                   num_blocks=len(dec.decoders))
    else:
        raise NotImplementedError("decoder class %s is outside the implemented set" % dname)
    cfg["model_conf"] = dict(ctc_weight=float(model.ctc_weight), reverse_weight=float(model.reverse_weight))
    return cfg


def wrap(model, device=None) -> B200ASRModel:
    """B200 core object sharing the weights of a loaded reference model."""
    core = B200ASRModel.from_reference(model, configs_from_reference_model(model), device=device)
    core.sos, core.eos = model.sos_symbol(), model.eos_symbol()
    return core


def install():
    """Rebind WeNet's registries (needs `wenet` importable).  Returns the plugin model class."""
    import wenet.dataset.processor as processor
    from wenet.models.transformer.asr_model import ASRModel
    from wenet.models.transformer.search import DecodeResult as RefDecodeResult
    from wenet.utils import init_model as im

    class B200ASRModelPlugin(ASRModel):
        """Reference ASRModel (same constructor, same parameters / state_dict) whose inference runs on
        libwenet_b200.so.  Training methods are inherited unchanged."""

        def _b200(self) -> B200ASRModel:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise _lib.WbError("the B200 path needs the model on a CUDA device (no CPU fallback); "
                                   "call model.to('cuda') first")
            core = getattr(self, "_b200_core", None)
            ver = sum(p._version for p in self.parameters())
            if core is None or getattr(self, "_b200_ver", None) != ver or core.device != dev:
                core = wrap(self, device=dev)
                object.__setattr__(self, "_b200_core", core)
                object.__setattr__(self, "_b200_ver", ver)
            return core

        def _forward_encoder(self, speech, speech_lengths, decoding_chunk_size=-1, num_decoding_left_chunks=-1,
                             simulate_streaming=False):
            return self._b200()._forward_encoder(speech, speech_lengths, decoding_chunk_size,
                                                 num_decoding_left_chunks, simulate_streaming)

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

    im.WENET_MODEL_CLASSES["asr_model"] = B200ASRModelPlugin
    processor.compute_fbank = b200_compute_fbank
    return B200ASRModelPlugin
