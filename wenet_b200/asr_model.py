"""Host-side mirror of the reference's model API for the hot path:

    B200ASRModel.decode(methods, speech, speech_lengths, beam_size, ...) -> Dict[str, List[DecodeResult]]
        == wenet/models/transformer/asr_model.py:267-343 (ASRModel.decode)
    B200ASRModel.encoder(xs, xs_lens, decoding_chunk_size, num_decoding_left_chunks) -> (xs, masks)
        == wenet/models/transformer/encoder.py:122-181
    B200ASRModel.encoder.forward_chunk(...) / forward_chunk_by_chunk(...)   == encoder.py:204-362
    B200ASRModel.ctc_logprobs / ctc.log_softmax                             == asr_model.py:254-265
    B200ASRModel.forward_attention_decoder(hyps, hyps_lens, encoder_out, reverse_weight)
        == asr_model.py:453-547

Same argument names, defaults and return types; every stage is a call into libwenet_b200.so with raw
device pointers (torch only allocates the buffers and supplies the stream).  No CPU fallback: all
tensors must live on a CUDA device.
"""
import ctypes as C
import math
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr
from .search import DecodeResult, LazyDecodeResult
from .weights import DeviceModel, ModelSpec


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class _EncOut:
    """Packed encoder output of one batch (rows of valid frames back to back)."""
    __slots__ = ("f32", "bf16", "seq_start", "seq_len", "lens_host", "starts_host", "rows", "max_len", "dump")


class _Subsampling:
    subsampling_rate = 4   # wenet/models/transformer/subsampling.py:196
    right_context = 6      # :197


class B200CTC:
    """model.ctc with the one method the decode path uses (ctc.py:73-81)."""

    def __init__(self, owner):
        self._o = owner

    def log_softmax(self, hs_pad: torch.Tensor) -> torch.Tensor:
        return self._o.ctc_logprobs(hs_pad)


class B200ConformerEncoder:
    """Drop-in for ConformerEncoder's inference methods."""

    def __init__(self, owner):
        self._o = owner
        self.embed = _Subsampling()
        spec = owner.spec
        self.use_dynamic_chunk = spec.use_dynamic_chunk
        self.static_chunk_size = spec.static_chunk_size

    def output_size(self) -> int:
        return self._o.spec.d_model

    # encoder.py:122-181
    def forward(self, xs: torch.Tensor, xs_lens: torch.Tensor, decoding_chunk_size: int = 0,
                num_decoding_left_chunks: int = -1) -> Tuple[torch.Tensor, torch.Tensor]:
        eo = self._o._encode(xs, xs_lens, decoding_chunk_size, num_decoding_left_chunks)
        return self._o._unpack(eo, xs.size(1))

    __call__ = forward

    def forward_chunk(self, xs, offset, required_cache_size, att_cache=None, cnn_cache=None, att_mask=None):
        return self._o._forward_chunk(xs, offset, required_cache_size, att_cache, cnn_cache)

    def forward_chunk_batch(self, xs, offsets, required_cache_size, att_cache=None, cnn_cache=None):
        """forward_chunk for S sessions in lockstep (batched caches, export_onnx_gpu.py:83-232): see
        B200ASRModel._forward_chunk_batch."""
        return self._o._forward_chunk_batch(xs, offsets, required_cache_size, att_cache, cnn_cache)

    def forward_chunk_by_chunk(self, xs: torch.Tensor, decoding_chunk_size: int,
                               num_decoding_left_chunks: int = -1) -> Tuple[torch.Tensor, torch.Tensor]:
        """Whole-utterance streaming simulation, same contract as BaseEncoder.forward_chunk_by_chunk
        (wenet/models/transformer/encoder.py:302-362): windows of (chunk-1)*4+7 frames every 4*chunk frames, the last
        one possibly shorter.  The steady-state steps are replayed as a CUDA graph by StreamingSession."""
        if decoding_chunk_size <= 0:
            raise ValueError("forward_chunk_by_chunk needs decoding_chunk_size > 0")
        if not (self.static_chunk_size > 0 or self.use_dynamic_chunk):
            raise ValueError("forward_chunk_by_chunk needs a model trained with chunk masks")
        sess = StreamingSession(self._o, decoding_chunk_size, num_decoding_left_chunks)
        hop = sess.window - self.embed.right_context - 1 + self.embed.subsampling_rate   # = 4 * chunk
        total = xs.size(1)
        pieces = [sess.step(xs[:, lo:lo + sess.window]).clone()
                  for lo in range(0, total - self.embed.right_context, hop)]
        ys = torch.cat(pieces, 1)
        return ys, torch.ones((1, 1, ys.size(1)), device=ys.device, dtype=torch.bool)


class StreamingSession:
    """Chunk-by-chunk streaming of ONE utterance through encoder.forward_chunk with the steady-state step replayed
    as a CUDA graph (the step is ~220 small launches: launch-bound on the host otherwise).

    Mirrors the loop of BaseEncoder.forward_chunk_by_chunk (encoder.py:302-362): call step(xs_chunk) with the
    (1, (chunk-1)*4+7, 80) feature window of each chunk; it returns y (1, chunk, d).  The first chunks (attention
    cache still growing) go through the regular call; once the cache has its final size the step is captured with
    static buffers and replayed - outputs are bit-identical to forward_chunk().
    """

    def __init__(self, model: "B200ASRModel", decoding_chunk_size: int, num_decoding_left_chunks: int,
                 use_graph: bool = True):
        assert decoding_chunk_size > 0
        self.m = model
        self.chunk = decoding_chunk_size
        self.required = decoding_chunk_size * num_decoding_left_chunks   # < 0: unbounded history (never steady)
        self.window = (decoding_chunk_size - 1) * 4 + 7
        self.use_graph = use_graph and self.required > 0 and model.spec.cnn_causal
        self.offset = 0
        dev = model.device
        self.att = torch.zeros(0, 0, 0, 0, device=dev)
        self.cnn = torch.zeros(0, 0, 0, 0, device=dev)
        self.graph = None
        self._warm = 0

    def _capture(self):
        m, spec, lib = self.m, self.m.spec, self.m._lib
        dev = m.device
        L, H, d, K = spec.enc_layers, spec.heads, spec.d_model, spec.cnn_kernel
        T, c1 = self.window, self.required
        self.s_xs = torch.zeros(T, spec.input_dim, device=dev, dtype=torch.float32)
        self.s_att = self.att.to(torch.float32).contiguous().clone()          # (L, H, c1, 128)
        self.s_cnn = self.cnn.to(torch.float32).contiguous().clone()          # (L, 1, d, K-1)
        self.s_ratt = torch.empty_like(self.s_att)
        self.s_rcnn = torch.empty_like(self.s_cnn)
        self.s_y = torch.empty(1, self.chunk, d, device=dev, dtype=torch.float32)
        self.s_off = torch.zeros(1, device=dev, dtype=torch.int32)   # position offset, advanced by the graph itself
        wsb = lib.wb_encoder_chunk_workspace_bytes(m.dm.handle, T, c1)
        self.s_ws = torch.empty(int(wsb) + 1024, device=dev, dtype=torch.uint8)
        oc, on = C.c_int(0), C.c_int(0)
        # a regular call leaves the shape block in the workspace (and warms every kernel's one-time attribute set-up)
        check(lib.wb_encoder_forward_chunk(m.dm.handle, ptr(self.s_xs), T, int(self.offset), int(self.required),
                                           ptr(self.s_att), c1, ptr(self.s_cnn), ptr(self.s_y), ptr(self.s_ratt),
                                           ptr(self.s_rcnn), C.byref(oc), C.byref(on), ptr(self.s_ws), self.s_ws.numel(),
                                           cur_stream()), "wb_encoder_forward_chunk")
        assert oc.value == self.chunk and on.value == c1
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            check(lib.wb_encoder_forward_chunk_static(m.dm.handle, ptr(self.s_xs), T, ptr(self.s_off), int(self.required),
                                                      ptr(self.s_att), c1, ptr(self.s_cnn), ptr(self.s_y),
                                                      ptr(self.s_ratt), ptr(self.s_rcnn), ptr(self.s_ws),
                                                      self.s_ws.numel(), cur_stream()), "wb_encoder_forward_chunk_static")
            self.s_att.copy_(self.s_ratt)     # the new caches are the next step's inputs
            self.s_cnn.copy_(self.s_rcnn)
            self.s_off.add_(self.chunk)
        self.s_off.fill_(self.offset)
        self.graph = g

    def step(self, xs: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(self.m.device):
            return self._step(xs)

    def _step(self, xs: torch.Tensor) -> torch.Tensor:
        m = self.m
        assert xs.size(0) == 1 and xs.size(1) <= self.window
        # only full windows with a full attention cache go through the graph; the first chunks and a shorter last
        # window (encoder.py:350 `end = min(cur + decoding_window, num_frames)`) take the regular call
        steady = (self.use_graph and xs.size(1) == self.window and self.att.numel() > 0
                  and self.att.size(2) == self.required)
        if steady and self.offset + self.chunk > m.spec.max_pos:
            raise _lib.WbError("utterance longer than the positional table")
        if not steady:
            if self.graph is not None:       # leaving the graph: its static buffers hold the current caches
                self.att, self.cnn = self.s_att, self.s_cnn
            y, self.att, self.cnn = m._forward_chunk(xs, self.offset, self.required, self.att, self.cnn)
            self.offset += y.size(1)
            if self.graph is not None:
                if self.att.shape == self.s_att.shape:
                    self.s_att.copy_(self.att)
                    self.s_cnn.copy_(self.cnn)
                self.s_off.fill_(self.offset)
            return y
        if self.graph is None:
            self._capture()
        self.s_xs.copy_(xs[0], non_blocking=True)
        self.graph.replay()
        self.offset += self.chunk
        return self.s_y            # static buffer: consume (or clone) before the next step


class BatchedStreamingSessions:
    """S concurrent streaming sessions advanced in lockstep (the batched caches of the reference's GPU
    export, wenet/bin/export_onnx_gpu.py:83-232): step(xs) takes the (S, (chunk-1)*4+7, idim) feature windows of all sessions
    and returns y (S, chunk, d); per session the arithmetic is forward_chunk's (encoder.py:204-300).  All sessions start at
    the same time (the attention caches grow together); once the caches have their final size the step is captured as a
    CUDA graph with static buffers (caches fed back, offsets advanced inside the graph)."""

    def __init__(self, model: "B200ASRModel", sessions: int, decoding_chunk_size: int, num_decoding_left_chunks: int,
                 use_graph: bool = True):
        assert decoding_chunk_size > 0 and sessions >= 1
        self.m, self.S = model, int(sessions)
        self.chunk = decoding_chunk_size
        self.required = decoding_chunk_size * num_decoding_left_chunks
        self.window = (decoding_chunk_size - 1) * 4 + 7
        self.use_graph = use_graph and self.required > 0 and model.spec.cnn_causal
        self.offsets = np.zeros(self.S, dtype=np.int32)
        self.att = None
        self.cnn = None
        self.graph = None

    def _capture(self):
        m, lib, S = self.m, self.m._lib, self.S
        dev, d = m.device, m.spec.d_model
        T, c1 = self.window, self.required
        self.s_xs = torch.zeros(S, T, m.spec.input_dim, device=dev, dtype=torch.float32)
        self.s_att, self.s_cnn = self.att.clone(), self.cnn.clone()
        self.s_ratt, self.s_rcnn = torch.empty_like(self.s_att), torch.empty_like(self.s_cnn)
        self.s_y = torch.empty(S, self.chunk, d, device=dev, dtype=torch.float32)
        self.s_off = torch.from_numpy(self.offsets.copy()).to(dev)
        wsb = lib.wb_encoder_chunk_batch_workspace_bytes(m.dm.handle, T, c1, S)
        self.s_ws = torch.empty(int(wsb) + 1024, device=dev, dtype=torch.uint8)
        oc, on = C.c_int(0), C.c_int(0)
        check(lib.wb_encoder_forward_chunk_batch(m.dm.handle, ptr(self.s_xs), T, S, ptr(_i32(self.offsets)), int(self.required),
                                                 ptr(self.s_att), c1, ptr(self.s_cnn), ptr(self.s_y), ptr(self.s_ratt),
                                                 ptr(self.s_rcnn), C.byref(oc), C.byref(on), ptr(self.s_ws),
                                                 self.s_ws.numel(), cur_stream()), "wb_encoder_forward_chunk_batch")
        assert oc.value == self.chunk and on.value == c1
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            check(lib.wb_encoder_forward_chunk_batch_static(m.dm.handle, ptr(self.s_xs), T, S, ptr(self.s_off),
                                                            int(self.required), ptr(self.s_att), c1, ptr(self.s_cnn),
                                                            ptr(self.s_y), ptr(self.s_ratt), ptr(self.s_rcnn), ptr(self.s_ws),
                                                            self.s_ws.numel(), cur_stream()),
                  "wb_encoder_forward_chunk_batch_static")
            self.s_att.copy_(self.s_ratt)
            self.s_cnn.copy_(self.s_rcnn)
            self.s_off.add_(self.chunk)
        self.graph = g

    def step(self, xs: torch.Tensor) -> torch.Tensor:
        m = self.m
        assert xs.size(0) == self.S
        with torch.cuda.device(m.device):
            steady = (self.use_graph and xs.size(1) == self.window and self.att is not None
                      and self.att.size(3) == self.required)
            if steady:
                if int(self.offsets.max()) + self.chunk > m.spec.max_pos:
                    raise _lib.WbError("utterance longer than the positional table")
                if self.graph is None:
                    self._capture()
                self.s_xs.copy_(xs, non_blocking=True)
                self.graph.replay()
                self.offsets += self.chunk
                return self.s_y        # static buffer: consume (or clone) before the next step
            if self.graph is not None:
                raise _lib.WbError("a window shorter than the steady-state one after graph capture: finish such sessions "
                                   "through forward_chunk / a fresh BatchedStreamingSessions")
            y, self.att, self.cnn = m._forward_chunk_batch(xs, self.offsets, self.required, self.att, self.cnn)
            self.offsets += y.size(1)
            return y


class B200ASRModel:
    """U2 / U2++ model (ConformerEncoder + CTC + (Bi)TransformerDecoder) on libwenet_b200.so."""

    def __init__(self, configs: dict, state_dict: Dict[str, torch.Tensor], device=None, with_decoder: bool = True,
                 precise: bool = False):
        """precise=True builds the PARITY mode (include/wenet_b200.h, wb_model_config.precise): bf16x3 encoder / CTC
        GEMMs + fp32 attention and depthwise conv; encoder_out and CTC log-probs within 1e-3 of the fp32 reference
        (tests/test_model_gpu.py).  The default is the bf16-operand throughput mode."""
        if not torch.cuda.is_available():
            raise _lib.WbError("B200ASRModel needs a CUDA device (there is no CPU fallback)")
        self.spec = ModelSpec(configs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            self.dm = DeviceModel(self.spec, state_dict, with_decoder=with_decoder, precise=precise)
        self.precise = bool(precise)
        self.vocab_size = self.spec.vocab
        self.sos = self.spec.sos            # asr_model.py:60-63 (tokenizer_conf.special_tokens or vocab - 1)
        self.eos = self.spec.eos
        self.ignore_id = -1
        self.reverse_weight = self.spec.reverse_weight
        self.ctc_weight = self.spec.ctc_weight
        self.default_decode_method = "attention_rescoring"
        self.encoder = B200ConformerEncoder(self)
        self.ctc = B200CTC(self)
        self._lib = _lib.load()
        self._ws = None
        self._ws2 = None
        self.keep_layer_dump = False
        self.d2h_bytes = 0      # bytes copied device -> host by decode() (bench.py: e2e.d2h_bytes_per_step)

    # ----- reference jit-export style accessors (asr_model.py:360-450) -----
    def subsampling_rate(self) -> int:
        return 4

    def right_context(self) -> int:
        return 6

    def sos_symbol(self) -> int:
        return self.sos

    def eos_symbol(self) -> int:
        return self.eos

    def is_bidirectional_decoder(self) -> bool:
        return self.spec.bidirectional and self.dm.has_decoder

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def clone_shared(self) -> "B200ASRModel":
        """A second front-end on the SAME device weights with its own workspaces — a finalized wb_model is
        immutable, so several host threads may decode concurrently on different streams (header comment
        of include/wenet_b200.h; mirrors TorchAsrModel::Copy, runtime/core/decoder/torch_asr_model.cc:87-111)."""
        import copy
        other = copy.copy(self)
        other.encoder = B200ConformerEncoder(other)
        other.ctc = B200CTC(other)
        other._ws = None
        other._ws2 = None
        other.d2h_bytes = 0
        other._pin = {}
        return other

    @classmethod
    def from_reference(cls, model, configs: dict, device=None, precise: bool = False):
        """Wrap a loaded reference ASRModel (same weights, same results, B200 kernels)."""
        return cls(configs, {k: v.detach().cpu() for k, v in model.state_dict().items()}, device=device,
                   precise=precise)

    def _operand(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 rows [R, d] -> the GEMM A operand the library expects for encoder output: bf16 [R, d], or in precise
        mode bf16 [R, 3d] = [hi | lo | hi]."""
        x = x.to(torch.float32).contiguous()
        R, d = x.shape
        p3 = 3 if self.precise else 1
        out = torch.empty(max(R, 1), d * p3, device=self.device, dtype=torch.bfloat16)
        if R > 0:
            check(self._lib.wb_op_cast_bf16(ptr(x), d, R, d, ptr(out), d * p3, int(self.precise), cur_stream()),
                  "wb_op_cast_bf16")
        return out

    def _host(self, t: torch.Tensor) -> np.ndarray:
        """device -> host copy of a result tensor (counted)."""
        self.d2h_bytes += t.numel() * t.element_size()
        return t.cpu().numpy()

    # ----- buffers -----
    def _workspace(self, nbytes: int, which: int = 0) -> torch.Tensor:
        cur = self._ws if which == 0 else self._ws2
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self.device)
            if which == 0:
                self._ws = cur
            else:
                self._ws2 = cur
        return cur

    # ----- encoder -----
    def _encode(self, speech: torch.Tensor, speech_lengths: torch.Tensor, decoding_chunk_size: int,
                num_decoding_left_chunks: int) -> _EncOut:
        if not speech.is_cuda:
            raise _lib.WbError("speech must be a CUDA tensor (no CPU fallback)")
        # add_optional_chunk_mask (wenet/utils/mask.py:162-198): the argument only matters for use_dynamic_chunk models
        if self.spec.use_dynamic_chunk:
            if decoding_chunk_size == 0:
                raise NotImplementedError("decoding_chunk_size=0 on a use_dynamic_chunk model selects the random "
                                          "*training* chunk (wenet/utils/mask.py:173-187); pass <0 (full) or >0")
            if decoding_chunk_size < 0:
                num_decoding_left_chunks = -1
        elif self.spec.static_chunk_size > 0:
            decoding_chunk_size = self.spec.static_chunk_size      # mask.py:188-194
        else:
            decoding_chunk_size, num_decoding_left_chunks = -1, -1  # mask.py:195-196: key-padding mask only
        x = speech.to(torch.float32).contiguous()
        B, T, D = x.shape
        assert D == self.spec.input_dim
        lens_host = _i32(speech_lengths.detach().cpu().numpy())
        assert int(lens_host.max()) <= T
        lib = self._lib
        rows = int(lib.wb_encoder_out_rows(B, ptr(lens_host)))
        d = self.spec.d_model
        eo = _EncOut()
        eo.rows = rows
        eo.f32 = torch.empty(max(rows, 1), d, device=self.device, dtype=torch.float32)
        eo.bf16 = torch.empty(max(rows, 1), d * (3 if self.precise else 1), device=self.device, dtype=torch.bfloat16)
        eo.seq_start = torch.zeros(B, device=self.device, dtype=torch.int32)
        eo.seq_len = torch.zeros(B, device=self.device, dtype=torch.int32)
        tp = np.where(lens_host >= 7, ((lens_host - 1) // 2 - 1) // 2, 0).astype(np.int32)
        eo.lens_host = tp
        eo.starts_host = np.concatenate([[0], np.cumsum(tp)[:-1]]).astype(np.int32)
        eo.max_len = int(tp.max()) if B else 0
        eo.dump = None
        if rows == 0:
            return eo
        wsb = lib.wb_encoder_workspace_bytes(self.dm.handle, B, ptr(lens_host))
        ws = self._workspace(wsb)
        dump = None
        if self.keep_layer_dump:
            dump = torch.empty(self.spec.enc_layers + 1, rows, d, device=self.device, dtype=torch.float32)
            eo.dump = dump
        pad_to = ((T - 1) // 2 - 1) // 2
        with torch.cuda.device(self.device):
            check(lib.wb_encoder_forward(self.dm.handle, ptr(x), x.stride(0), ptr(lens_host), B,
                                         int(decoding_chunk_size), int(num_decoding_left_chunks), int(pad_to),
                                         ptr(eo.f32), ptr(eo.bf16), ptr(eo.seq_start), ptr(eo.seq_len), ptr(dump),
                                         ptr(ws), ws.numel(), cur_stream()), "wb_encoder_forward")
        return eo

    def _unpack(self, eo: _EncOut, T_in: int) -> Tuple[torch.Tensor, torch.Tensor]:
        B = eo.seq_start.numel()
        Tp = ((T_in - 1) // 2 - 1) // 2 if T_in >= 7 else 0
        d = self.spec.d_model
        out = torch.empty(B, Tp, d, device=self.device, dtype=torch.float32)
        if Tp > 0:
            check(self._lib.wb_unpack_rows(ptr(eo.f32), ptr(eo.seq_start), ptr(eo.seq_len), B, Tp, d, ptr(out), Tp,
                                           cur_stream()), "wb_unpack_rows")
        masks = (torch.arange(Tp, device=self.device).unsqueeze(0) < eo.seq_len.unsqueeze(1)).unsqueeze(1)
        return out, masks

    def _forward_encoder(self, speech, speech_lengths, decoding_chunk_size=-1, num_decoding_left_chunks=-1,
                         simulate_streaming=False):
        # asr_model.py:216-239
        if simulate_streaming and decoding_chunk_size > 0:
            return self.encoder.forward_chunk_by_chunk(speech, decoding_chunk_size, num_decoding_left_chunks)
        return self.encoder.forward(speech, speech_lengths, decoding_chunk_size, num_decoding_left_chunks)

    def _forward_chunk(self, xs, offset, required_cache_size, att_cache, cnn_cache):
        if not xs.is_cuda:
            raise _lib.WbError("xs must be a CUDA tensor (no CPU fallback)")
        assert xs.size(0) == 1
        spec = self.spec
        lib = self._lib
        x = xs[0].to(torch.float32).contiguous()
        T = x.size(0)
        L, H, d, K = spec.enc_layers, spec.heads, spec.d_model, spec.cnn_kernel
        cache_t1 = att_cache.size(2) if (att_cache is not None and att_cache.numel() > 0) else 0
        chunk = ((T - 1) // 2 - 1) // 2
        key_size = cache_t1 + chunk
        if required_cache_size < 0:
            nxt = 0
        elif required_cache_size == 0:
            nxt = key_size
        else:
            nxt = max(key_size - required_cache_size, 0)
        y = torch.empty(1, chunk, d, device=self.device, dtype=torch.float32)
        r_att = torch.empty(L, H, key_size - nxt, 128, device=self.device, dtype=torch.float32)
        r_cnn = torch.empty(L, 1, d, max(K - 1, 0), device=self.device, dtype=torch.float32) if spec.cnn_causal \
            else torch.zeros(0, 0, 0, 0, device=self.device)
        ac = att_cache.to(torch.float32).contiguous() if cache_t1 > 0 else None
        cc = cnn_cache.to(torch.float32).contiguous() if (cnn_cache is not None and cnn_cache.numel() > 0) else None
        wsb = lib.wb_encoder_chunk_workspace_bytes(self.dm.handle, T, cache_t1)
        ws = self._workspace(wsb)
        oc, on = C.c_int(0), C.c_int(0)
        with torch.cuda.device(self.device):
            check(lib.wb_encoder_forward_chunk(self.dm.handle, ptr(x), T, int(offset), int(required_cache_size), ptr(ac),
                                               cache_t1, ptr(cc), ptr(y), ptr(r_att),
                                               ptr(r_cnn) if spec.cnn_causal else None, C.byref(oc), C.byref(on),
                                               ptr(ws), ws.numel(), cur_stream()), "wb_encoder_forward_chunk")
        return y, r_att, r_cnn

    def _forward_chunk_batch(self, xs, offsets, required_cache_size, att_cache, cnn_cache):
        """S sessions in lockstep: xs (S, T, idim); offsets (S,) ints; att_cache (S, L, H, cache_t1, 128) or None;
        cnn_cache (S, L, d, K-1) or None.  Returns (y (S, chunk, d), r_att_cache, r_cnn_cache) with the same layouts."""
        if not xs.is_cuda:
            raise _lib.WbError("xs must be a CUDA tensor (no CPU fallback)")
        spec, lib = self.spec, self._lib
        x = xs.to(torch.float32).contiguous()
        S, T, _ = x.shape
        L, H, d, K = spec.enc_layers, spec.heads, spec.d_model, spec.cnn_kernel
        cache_t1 = att_cache.size(3) if (att_cache is not None and att_cache.numel() > 0) else 0
        chunk = ((T - 1) // 2 - 1) // 2
        key_size = cache_t1 + chunk
        if required_cache_size < 0:
            nxt = 0
        elif required_cache_size == 0:
            nxt = key_size
        else:
            nxt = max(key_size - required_cache_size, 0)
        lead = K - 1 if spec.cnn_causal else 0
        y = torch.empty(S, chunk, d, device=self.device, dtype=torch.float32)
        r_att = torch.empty(S, L, H, key_size - nxt, 128, device=self.device, dtype=torch.float32)
        r_cnn = torch.empty(S, L, d, max(lead, 0), device=self.device, dtype=torch.float32)
        ac = att_cache.to(torch.float32).contiguous() if cache_t1 > 0 else None
        cc = cnn_cache.to(torch.float32).contiguous() if (cnn_cache is not None and cnn_cache.numel() > 0) else None
        wsb = lib.wb_encoder_chunk_batch_workspace_bytes(self.dm.handle, T, cache_t1, S)
        ws = self._workspace(wsb)
        oc, on = C.c_int(0), C.c_int(0)
        with torch.cuda.device(self.device):
            check(lib.wb_encoder_forward_chunk_batch(self.dm.handle, ptr(x), T, S, ptr(_i32(offsets)), int(required_cache_size),
                                                     ptr(ac), cache_t1, ptr(cc), ptr(y), ptr(r_att),
                                                     ptr(r_cnn) if lead > 0 else None, C.byref(oc), C.byref(on), ptr(ws),
                                                     ws.numel(), cur_stream()), "wb_encoder_forward_chunk_batch")
        return y, r_att, r_cnn

    # ----- CTC -----
    def _ctc(self, eo: _EncOut, topk: int, blank_id: int, blank_penalty: float, full: bool = True):
        """full=True: normalised log-probs for every token (API parity); False: only the per-frame top-k (decode())."""
        V = self.spec.vocab
        ldl = (V + 7) // 8 * 8
        rows = max(eo.rows, 1)
        logp = torch.empty(rows, ldl, device=self.device, dtype=torch.float32)
        k = max(int(topk), 1)
        tv = torch.empty(rows, k, device=self.device, dtype=torch.float32)
        ti = torch.empty(rows, k, device=self.device, dtype=torch.int32)
        if eo.rows > 0:
            fn = self._lib.wb_ctc_logprobs if (full or k > 64) else self._lib.wb_ctc_topk
            check(fn(self.dm.handle, ptr(eo.bf16), eo.rows, int(blank_id), float(blank_penalty), ptr(logp), ldl, k, ptr(tv),
                     ptr(ti), cur_stream()), "wb_ctc_logprobs")
        return logp, tv, ti

    def ctc_logprobs(self, encoder_out: torch.Tensor, blank_penalty: float = 0.0, blank_id: int = 0) -> torch.Tensor:
        """asr_model.py:254-265 on a padded (B, T', d) tensor (API parity; decode() uses the packed path)."""
        if not encoder_out.is_cuda:
            raise _lib.WbError("encoder_out must be a CUDA tensor (no CPU fallback)")
        B, Tp, d = encoder_out.shape
        eo = _EncOut()
        eo.rows = B * Tp
        with torch.cuda.device(self.device):
            eo.bf16 = self._operand(encoder_out.reshape(B * Tp, d))
            logp, _, _ = self._ctc(eo, 1, blank_id, blank_penalty)
        return logp[:, :self.spec.vocab].reshape(B, Tp, self.spec.vocab)

    # ----- decode (asr_model.py:267-343) -----
    def decode(self, methods: List[str], speech: torch.Tensor, speech_lengths: torch.Tensor, beam_size: int = 1,
               decoding_chunk_size: int = -1, num_decoding_left_chunks: int = -1, ctc_weight: float = 0.0,
               simulate_streaming: bool = False, reverse_weight: float = 0.0, context_graph=None,
               blank_id: int = 0, blank_penalty: float = 0.0, length_penalty: float = 0.0,
               infos: Dict[str, List[str]] = None) -> Dict[str, List[DecodeResult]]:
        assert speech.shape[0] == speech_lengths.shape[0]
        assert decoding_chunk_size != 0
        cg = None
        if context_graph is not None:
            from . import context as _ctx
            cg = _ctx.flatten(context_graph)      # the reference's ContextGraph object, or ContextArrays
        with torch.cuda.device(self.device):
            if simulate_streaming and decoding_chunk_size > 0:
                ys, _ = self.encoder.forward_chunk_by_chunk(speech, decoding_chunk_size, num_decoding_left_chunks)
                eo = self._pack_padded(ys)
            else:
                eo = self._encode(speech, speech_lengths, decoding_chunk_size, num_decoding_left_chunks)
            need_beam = ("ctc_prefix_beam_search" in methods) or ("attention_rescoring" in methods)
            topk = beam_size if need_beam else 1
            results = {}
            if "attention" in methods:
                # asr_model.py:315-318 -> search.py:252-371; maxlen = encoder_out.size(1) of the padded batch
                T_in = speech.size(1)
                maxlen = (((T_in - 1) // 2 - 1) // 2 if T_in >= 7 else 0) if not (simulate_streaming and decoding_chunk_size > 0) \
                    else eo.max_len
                prefix = np.full((speech.shape[0], 1), self.sos, dtype=np.int32)
                results["attention"] = self._attention_beam(eo, beam_size, length_penalty, prefix, self.eos, maxlen)
                if len(methods) == 1:
                    return results
            logp, tv, ti = self._ctc(eo, topk, blank_id, blank_penalty, full=False)
            if "ctc_greedy_search" in methods:
                results["ctc_greedy_search"] = self._greedy(eo, ti, blank_id)
            if need_beam:
                # device-resident pipeline: only the per-hypothesis lengths / scores cross to the host before the
                # rescoring pass is enqueued; the Python result objects are built while the decoder runs
                bd = self._prefix_beam_launch(eo, tv, ti, beam_size, blank_id, cg)
                meta = self._beam_meta(bd)
                fetch = self._beam_fetch_async(bd, meta)
                rs = None
                if "attention_rescoring" in methods:
                    rs = self._rescore_launch(eo, bd, meta, fetch, ctc_weight, reverse_weight)
                beam_out = self._beam_results(bd, meta, fetch)
                if "ctc_prefix_beam_search" in methods:
                    results["ctc_prefix_beam_search"] = beam_out
                if rs is not None:
                    results["attention_rescoring"] = self._rescore_results(rs, meta, beam_out, reverse_weight)
        return results

    def _attention_beam(self, eo: _EncOut, beam_size: int, length_penalty: float, prefix: np.ndarray, eos: int,
                        maxlen: int) -> List[DecodeResult]:
        """attention_beam_search (search.py:252-371) on the library: prefix [B, P] forced start tokens, maxlen =
        encoder_out.size(1) of the padded batch (the reference's step bound)."""
        if not self.dm.has_decoder:
            raise _lib.WbError("decode mode 'attention' needs a decoder")
        lib = self._lib
        B, P = prefix.shape
        max_tok = int(maxlen) + 1
        if eo.rows == 0 or max_tok <= P:
            return [DecodeResult([]) for _ in range(B)]
        pe_len = self.spec.dec_max_len if self.spec.dec_flavor == 1 else self.spec.max_pos
        max_tok = min(max_tok, pe_len + 1)     # the decoder's position table bounds the hypothesis length
        stride = max_tok - P
        toks = torch.zeros(B, stride, device=self.device, dtype=torch.int32)
        lens = torch.zeros(B, device=self.device, dtype=torch.int32)
        scores = torch.zeros(B, device=self.device, dtype=torch.float32)
        wsb = lib.wb_attention_beam_workspace_bytes(self.dm.handle, eo.rows, B, int(beam_size), max_tok)
        ws = self._workspace(wsb, 1)
        steps = C.c_int32(0)
        check(lib.wb_attention_beam_search(self.dm.handle, ptr(eo.bf16), eo.rows, ptr(_i32(eo.starts_host)),
                                           ptr(_i32(eo.lens_host)), B, int(beam_size), ptr(_i32(prefix)), P, int(eos), max_tok,
                                           float(length_penalty), ptr(toks), stride, ptr(lens), ptr(scores), C.byref(steps),
                                           ptr(ws), wsb, cur_stream()), "wb_attention_beam_search")
        self.last_attention_steps = int(steps.value)
        th, lh = self._host(toks), self._host(lens)
        return [DecodeResult(th[b, :lh[b]].tolist()) for b in range(B)]

    def _pack_padded(self, ys: torch.Tensor) -> _EncOut:
        B, Tp, d = ys.shape
        eo = _EncOut()
        eo.rows = B * Tp
        eo.f32 = ys.reshape(B * Tp, d).contiguous()
        eo.bf16 = self._operand(eo.f32)
        eo.lens_host = np.full(B, Tp, dtype=np.int32)
        eo.starts_host = (np.arange(B) * Tp).astype(np.int32)
        eo.seq_start = torch.from_numpy(eo.starts_host).to(self.device)
        eo.seq_len = torch.from_numpy(eo.lens_host).to(self.device)
        eo.max_len = Tp
        eo.dump = None
        return eo

    def _greedy(self, eo: _EncOut, ti: torch.Tensor, blank_id: int) -> List[DecodeResult]:
        B = eo.seq_start.numel()
        stride = max(eo.max_len, 1)
        toks = torch.zeros(B, stride, device=self.device, dtype=torch.int32)
        lens = torch.zeros(B, device=self.device, dtype=torch.int32)
        check(self._lib.wb_ctc_greedy_search(ptr(ti), ti.stride(0), ptr(eo.seq_start), ptr(eo.seq_len), B, int(blank_id),
                                             ptr(toks), stride, ptr(lens), cur_stream()), "wb_ctc_greedy_search")
        th, lh = self._host(toks), self._host(lens)
        return [DecodeResult(th[b, :lh[b]].tolist()) for b in range(B)]

    # pinned staging buffers for asynchronous device -> host copies (allocated once, grown on demand)
    def _pinned(self, name: str, numel: int, dtype) -> torch.Tensor:
        bufs = self.__dict__.setdefault("_pin", {})
        cur = bufs.get(name)
        if cur is None or cur.numel() < numel or cur.dtype != dtype:
            cur = torch.empty(max(int(numel * 1.25), 1024), dtype=dtype, pin_memory=True)
            bufs[name] = cur
        return cur[:numel]

    def _d2h_async(self, name: str, t: torch.Tensor) -> np.ndarray:
        """enqueue a device -> pinned-host copy on the current stream; valid after the next event/stream sync"""
        t = t.contiguous()
        dst = self._pinned(name, t.numel(), t.dtype)
        dst.copy_(t.view(-1), non_blocking=True)
        self.d2h_bytes += t.numel() * t.element_size()
        return dst.numpy().reshape(tuple(t.shape))

    def _prefix_beam_launch(self, eo: _EncOut, tv, ti, beam_size: int, blank_id: int, cg=None):
        B = eo.seq_start.numel()
        max_len = max(eo.max_len, 1)
        lib = self._lib
        bd = SimpleNamespace()
        bd.B, bd.beam, bd.max_len = B, beam_size, max_len
        bd.toks = torch.zeros(B, beam_size, max_len, device=self.device, dtype=torch.int32)
        bd.times = torch.zeros(B, beam_size, max_len, device=self.device, dtype=torch.int32)
        bd.lens = torch.zeros(B, beam_size, device=self.device, dtype=torch.int32)
        bd.scores = torch.zeros(B, beam_size, device=self.device, dtype=torch.float64)
        bd.nhyp = torch.zeros(B, device=self.device, dtype=torch.int32)
        wsb = lib.wb_prefix_beam_workspace_bytes(B, beam_size, max_len)
        ws = self._workspace(wsb, 1)
        cgs = None
        if cg is not None:
            from . import context as _ctx
            cgs = C.byref(_ctx.to_device(cg, self.device))
        check(lib.wb_ctc_prefix_beam_search_ctx(ptr(tv), ptr(ti), tv.stride(0), ptr(eo.seq_start), ptr(eo.seq_len), B,
                                                int(beam_size), int(blank_id), max_len, cgs, ptr(bd.toks), ptr(bd.times),
                                                ptr(bd.lens), ptr(bd.scores), ptr(bd.nhyp), ptr(ws), wsb, cur_stream()),
              "wb_ctc_prefix_beam_search")
        return bd

    def _beam_meta(self, bd):
        """lengths / scores / counts of the n-best lists -> flat utterance-major hypothesis tables (host, small)"""
        lh = self._d2h_async("beam_lens", bd.lens)
        sh = self._d2h_async("beam_scores", bd.scores)
        nh = self._d2h_async("beam_nhyp", bd.nhyp)
        torch.cuda.current_stream().synchronize()
        m = SimpleNamespace()
        valid = np.arange(bd.beam)[None, :] < nh[:, None]                          # (B, beam)
        bi, ri = np.nonzero(valid)
        m.nh = nh.copy()
        m.hyp_utt = bi.astype(np.int32)
        m.hyp_len = np.ascontiguousarray(lh[valid].astype(np.int32))
        m.hyp_src = ((bi * bd.beam + ri) * bd.max_len).astype(np.int32)             # offsets into bd.toks (device)
        m.scores = np.ascontiguousarray(sh[valid].astype(np.float64))
        m.slot = (bi * bd.beam + ri).astype(np.int64)
        m.lmax = int(m.hyp_len.max()) if m.hyp_len.size else 0
        return m

    def _beam_fetch_async(self, bd, meta):
        L = max(meta.lmax, 1)
        th = self._d2h_async("beam_toks", bd.toks[:, :, :L])
        mh = self._d2h_async("beam_times", bd.times[:, :, :L])
        ev = torch.cuda.Event()
        ev.record()
        return th, mh, ev

    def _beam_results(self, bd, meta, fetch) -> List[DecodeResult]:
        th, mh, ev = fetch
        ev.synchronize()
        L = th.shape[2]
        th = th.reshape(-1, L)[meta.slot]            # (n_hyp, L) copies: the pinned staging is reused by the next decode
        mh = mh.reshape(-1, L)[meta.slot]
        score_list = meta.scores.tolist()
        ends = np.cumsum(meta.nh).tolist()
        out = []
        h = 0
        for b in range(bd.B):
            e = ends[b]
            out.append(LazyDecodeResult(score_list[h], score_list[h:e], th[h:e], mh[h:e], meta.hyp_len[h:e]))
            h = e
        return out

    def _flatten_hyps(self, hyps_per_utt):
        hyp_utt, hyp_len, hyp_tok0, toks = [], [], [], []
        for b, hyps in enumerate(hyps_per_utt):
            for h in hyps:
                hyp_utt.append(b)
                hyp_len.append(len(h))
                hyp_tok0.append(len(toks))
                toks.extend(int(t) for t in h)
        if not toks:
            toks = [0]
        return _i32(hyp_utt), _i32(hyp_len), _i32(hyp_tok0), _i32(toks)

    def _rescore_launch(self, eo: _EncOut, bd, meta, fetch, ctc_weight: float, reverse_weight: float):
        """search.py:374-458.  The n-best token ids (a few hundred KB, already on their way to pinned host memory for the
        result objects) are handed to the library on the host side, which lets it share decoder rows between
        hypotheses with a common prefix; WB_RESCORE_DEVICE_TOKENS=1 keeps them on the device (no sharing)."""
        if not self.dm.has_decoder:
            raise _lib.WbError("attention_rescoring needs decoder weights")
        lib = self._lib
        B = bd.B
        n_hyp = int(meta.hyp_utt.size)
        R = int(meta.hyp_len.sum()) + n_hyp
        rs = SimpleNamespace()
        rs.use_r2l = reverse_weight > 0 and self.spec.bidirectional and self.spec.rdec_layers > 0
        l2r = torch.zeros(R, device=self.device, dtype=torch.float32)
        r2l = torch.zeros(R, device=self.device, dtype=torch.float32)
        hyp_score = torch.zeros(n_hyp, device=self.device, dtype=torch.float32)
        best = torch.zeros(B, device=self.device, dtype=torch.int32)
        wsb = lib.wb_rescoring_workspace_bytes(self.dm.handle, eo.rows, R)
        ws = self._workspace(wsb)
        if os.environ.get("WB_RESCORE_DEVICE_TOKENS"):
            check(lib.wb_attention_rescoring_dev(self.dm.handle, ptr(eo.bf16), eo.rows, ptr(eo.starts_host),
                                                 ptr(eo.lens_host), B, n_hyp, ptr(meta.hyp_utt), ptr(meta.hyp_len),
                                                 ptr(meta.hyp_src), ptr(bd.toks), ptr(meta.scores), self.sos, self.eos,
                                                 float(ctc_weight), float(reverse_weight if rs.use_r2l else 0.0), ptr(l2r),
                                                 ptr(r2l), ptr(hyp_score), ptr(best), ptr(ws), ws.numel(), cur_stream()),
                  "wb_attention_rescoring_dev")
        else:
            th, _, ev = fetch
            ev.synchronize()                                   # token ids have landed in pinned host memory
            tok0 = np.ascontiguousarray((meta.slot * th.shape[2]).astype(np.int32))   # hyp h starts at th[b, rank, 0]
            check(lib.wb_attention_rescoring(self.dm.handle, ptr(eo.bf16), eo.rows, ptr(eo.starts_host),
                                             ptr(eo.lens_host), B, n_hyp, ptr(meta.hyp_utt), ptr(meta.hyp_len),
                                             ptr(tok0), ptr(th), ptr(meta.scores), self.sos, self.eos,
                                             float(ctc_weight), float(reverse_weight if rs.use_r2l else 0.0), ptr(l2r),
                                             ptr(r2l), ptr(hyp_score), ptr(best), ptr(ws), ws.numel(), cur_stream()),
                  "wb_attention_rescoring")
        rs.l2r = self._d2h_async("rs_l2r", l2r)
        rs.r2l = self._d2h_async("rs_r2l", r2l) if rs.use_r2l else None
        rs.hs = self._d2h_async("rs_score", hyp_score)
        rs.best = self._d2h_async("rs_best", best)
        rs.ev = torch.cuda.Event()
        rs.ev.record()
        return rs

    def _rescore_results(self, rs, meta, beam_out: List[DecodeResult], reverse_weight: float):
        rs.ev.synchronize()
        B = len(beam_out)
        use_r2l = rs.use_r2l
        hyp_len = meta.hyp_len
        hs = rs.hs.copy()                        # pinned staging buffers are reused by the next decode
        bh = rs.best.tolist()
        row0 = np.concatenate([[0], np.cumsum(hyp_len.astype(np.int64) + 1)])
        ends = np.cumsum(meta.nh).tolist()
        hs_list = hs.tolist()
        out = []
        h0 = 0
        for b in range(B):
            bi = bh[b]
            h = h0 + bi
            n = int(hyp_len[h])
            r0 = int(row0[h])
            seg = rs.l2r[r0:r0 + n + 1].copy()
            rseg = rs.r2l[r0:r0 + n + 1].copy() if use_r2l else None

            def conf_fn(seg=seg, rseg=rseg, n=n):
                tc = np.exp(seg[:n].astype(np.float64))
                score = np.cumsum(seg, dtype=np.float32)[-1]          # sequential fp32 sum, as the reference
                if rseg is not None:
                    # r_decoder_out[i][len-j-1][hyp[j]] (search.py:438-441): position n-1-j scores token j
                    tc = (tc + np.exp(rseg[:n][::-1].astype(np.float64))) / 2
                    r_score = np.cumsum(np.concatenate([rseg[:n][::-1], rseg[n:n + 1]]), dtype=np.float32)[-1]
                    score = np.float32(score * np.float32(1 - reverse_weight) + r_score * np.float32(reverse_weight))
                return math.exp(float(score) / (n + 1)), tc.tolist()

            bo = beam_out[b]
            # tokens / times of the chosen hypothesis come from the beam result's packed rows; nbest_scores carries all
            # rescored hypotheses of the utterance (extra over the reference, which returns only the best)
            out.append(LazyDecodeResult(hs_list[h], hs_list[h0:ends[b]], bo._tok_rows, bo._time_rows, bo._lens, best=bi,
                                        conf_fn=conf_fn))
            h0 = ends[b]
        return out

    # ----- asr_model.py:453-547 -----
    def forward_attention_decoder(self, hyps: torch.Tensor, hyps_lens: torch.Tensor, encoder_out: torch.Tensor,
                                  reverse_weight: float = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """hyps (N, L) with <sos> first, hyps_lens (N,) counting <sos>; encoder_out (1, T', d).
        Returns (log-probs (N, L, V), r_log-probs (N, L, V) or tensor(0.)).  Positions past a
        hypothesis' own length are zero-filled (the reference computes don't-care values there)."""
        if not encoder_out.is_cuda:
            raise _lib.WbError("encoder_out must be a CUDA tensor (no CPU fallback)")
        assert encoder_out.size(0) == 1
        lib = self._lib
        N, L = hyps.shape
        V = self.spec.vocab
        lens = (hyps_lens.detach().cpu().numpy().astype(np.int64) - 1)
        hy = hyps.detach().cpu().numpy()
        hyp_list = [hy[i, 1:1 + lens[i]].tolist() for i in range(N)]
        hyp_utt, hyp_len, hyp_tok0, toks = self._flatten_hyps([hyp_list])
        Tp = encoder_out.size(1)
        enc_bf16 = self._operand(encoder_out[0])
        R = int(hyp_len.sum()) + N
        ldl = (V + 7) // 8 * 8
        use_r2l = reverse_weight > 0 and self.is_bidirectional_decoder()
        lp = torch.empty(R, ldl, device=self.device, dtype=torch.float32)
        rlp = torch.empty(R, ldl, device=self.device, dtype=torch.float32) if use_r2l else None
        wsb = lib.wb_rescoring_workspace_bytes(self.dm.handle, Tp, R)
        ws = self._workspace(wsb)
        starts, slens = _i32([0]), _i32([Tp])   # keep the host arrays alive across the call
        with torch.cuda.device(self.device):
            check(lib.wb_decoder_logprobs(self.dm.handle, ptr(enc_bf16), Tp, ptr(starts), ptr(slens), 1, N,
                                          ptr(hyp_utt), ptr(hyp_len), ptr(hyp_tok0), ptr(toks), int(hy[0, 0]), self.eos,
                                          int(use_r2l), ptr(lp), ptr(rlp), ldl, ptr(ws), ws.numel(), cur_stream()),
                  "wb_decoder_logprobs")
        out = torch.zeros(N, L, V, device=self.device, dtype=torch.float32)
        r_out = torch.zeros(N, L, V, device=self.device, dtype=torch.float32) if use_r2l else torch.tensor(0.0)
        r0 = 0
        for i in range(N):
            n = int(hyp_len[i]) + 1
            out[i, :n] = lp[r0:r0 + n, :V]
            if use_r2l:
                r_out[i, :n] = rlp[r0:r0 + n, :V]
            r0 += n
        return out, r_out
