"""Whisper on libwenet_b200.so (BASELINE configs[4]): host-side mirror of

    wenet/models/whisper/whisper.py:28-96      Whisper(ASRModel): sos = sot, eos = eot, default mode "attention"
    wenet/dataset/processor.py:320-369         compute_log_mel_spectrogram
    wenet/models/transformer/encoder.py:365-440 TransformerEncoder (conv1d2 / abs_pos_whisper / gelu) .forward
    wenet/models/transformer/search.py:252-371 attention_beam_search with the Whisper prefix
    wenet/utils/common.py:159-238              add_whisper_tokens (the forced start [sot, language, task, no_timestamps])

Same call signatures and result types as the reference; every stage is a C-ABI call (include/wenet_b200.h sections F, G).
No CPU fallback.
"""
import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr
from .asr_model import B200ASRModel, B200CTC, _EncOut, _i32
from .search import DecodeResult

# openai-whisper tokenizer.LANGUAGES key order (wenet/utils/common.py:24-26 WHISPER_LANGS); language id = sot + 1 + index
WHISPER_LANGS = (
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
    "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
    "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
    "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue")


def slaney_mel_filters(sr: int = 16000, n_fft: int = 400, n_mels: int = 128) -> np.ndarray:
    """librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels) (htk=False, norm='slaney', fmin 0, fmax sr/2), restated
    from its published algorithm - the reference calls it at processor.py:360-361; librosa is not installed here.
    Returns float32 [n_mels, n_fft // 2 + 1]."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        freqs = f_sp * m
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)

    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(float(sr) / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


class LogMelExtractor:
    """Batched on-device compute_log_mel_spectrogram (processor.py:320-369)."""

    def __init__(self, num_mel_bins: int = 80, n_fft: int = 400, hop_length: int = 160, sample_rate: int = 16000):
        self.num_mel, self.n_fft, self.hop = int(num_mel_bins), int(n_fft), int(hop_length)
        lib = _lib.load()
        win = torch.hann_window(self.n_fft).numpy().astype(np.float32)
        mel = np.ascontiguousarray(slaney_mel_filters(sample_rate, self.n_fft, self.num_mel))
        self._h = C.c_void_p()
        check(lib.wb_logmel_create(C.byref(self._h), self.n_fft, self.hop, self.num_mel, ptr(np.ascontiguousarray(win)), ptr(mel)),
              "wb_logmel_create")
        self._lib = lib

    def num_frames(self, n: int) -> int:
        return n // self.hop

    def __call__(self, pcm: torch.Tensor, num_samples: torch.Tensor, max_frames: Optional[int] = None) -> torch.Tensor:
        """pcm [B, S] float32 in [-1, 1) on the GPU, num_samples [B] int32 (device) -> feats [B, max_frames, num_mel]"""
        if not pcm.is_cuda:
            raise _lib.WbError("pcm must be a CUDA tensor (no CPU fallback)")
        pcm = pcm.to(torch.float32).contiguous()
        B, S = pcm.shape
        ns = num_samples.to(device=pcm.device, dtype=torch.int32).contiguous()
        if max_frames is None:
            max_frames = S // self.hop
        out = torch.empty(B, max_frames, self.num_mel, device=pcm.device, dtype=torch.float32)
        scratch = torch.empty(B, device=pcm.device, dtype=torch.int32)
        with torch.cuda.device(pcm.device):
            check(self._lib.wb_logmel_forward(self._h, ptr(pcm), pcm.stride(0), ptr(ns), B, ptr(out), max_frames, max_frames,
                                              ptr(scratch), cur_stream()), "wb_logmel_forward")
        return out

    def __del__(self):
        try:
            if self._h:
                self._lib.wb_logmel_destroy(self._h)
        except Exception:
            pass


_extractors: Dict[Tuple[int, int, int], LogMelExtractor] = {}


def compute_log_mel_spectrogram(sample, n_fft=400, hop_length=160, num_mel_bins=80, padding=0, pad_or_trim: bool = False,
                                max_duration: int = 30):
    """Drop-in for wenet.dataset.processor.compute_log_mel_spectrogram (processor.py:320-369): same sample dict in / out."""
    assert 'sample_rate' in sample and 'wav' in sample and 'key' in sample
    sr = sample['sample_rate']
    wav = sample['wav'].squeeze(0)
    if padding > 0:
        wav = torch.nn.functional.pad(wav, (0, padding))
    if pad_or_trim:
        length = max_duration * sr
        wav = wav[:length] if wav.size(0) >= length else torch.nn.functional.pad(wav, (0, length - wav.size(0)))
    key = (int(num_mel_bins), int(n_fft), int(hop_length))
    ex = _extractors.get(key)
    if ex is None:
        ex = _extractors[key] = LogMelExtractor(num_mel_bins, n_fft, hop_length, sr)
    dev = torch.device("cuda", torch.cuda.current_device())
    n = wav.size(0)
    feats = ex(wav.to(dev).unsqueeze(0), torch.tensor([n], dtype=torch.int32, device=dev))
    sample['feat'] = feats[0].cpu()
    return sample


def whisper_prefix(special_tokens: dict, tasks: List[str], langs: List[str], lang_table=WHISPER_LANGS) -> np.ndarray:
    """add_whisper_tokens(..., no_timestamp=True, use_prev=False) on empty hypotheses (common.py:198-226): one forced
    start [sot, language, task, no_timestamps | no_speech] per utterance."""
    rows = []
    for task, lang in zip(tasks, langs):
        if task == "transcribe":
            task_id = special_tokens["transcribe"]
        elif task == "translate":
            task_id = special_tokens["translate"]
        elif task == "vad":
            task_id = special_tokens["no_speech"]
        else:
            raise NotImplementedError("unsupported task {}".format(task))
        language_id = special_tokens["sot"] + 1 + list(lang_table).index(lang)
        prefix = [special_tokens["sot"], language_id, task_id]
        prefix.append(special_tokens["no_timestamps"] if task in ("transcribe", "translate") else special_tokens["no_speech"])
        rows.append(prefix)
    return np.asarray(rows, dtype=np.int32)


class _Conv1dSubsampling2:
    subsampling_rate = 2   # subsampling.py:138
    right_context = 4      # :140


class B200TransformerEncoder:
    """Drop-in for the inference `forward` of the Whisper TransformerEncoder (encoder.py:122-181)."""

    def __init__(self, owner):
        self._o = owner
        self.embed = _Conv1dSubsampling2()
        self.use_dynamic_chunk = False
        self.static_chunk_size = 0

    def output_size(self) -> int:
        return self._o.spec.d_model

    def forward(self, xs: torch.Tensor, xs_lens: torch.Tensor, decoding_chunk_size: int = 0,
                num_decoding_left_chunks: int = -1) -> Tuple[torch.Tensor, torch.Tensor]:
        eo = self._o._encode(xs, xs_lens, decoding_chunk_size, num_decoding_left_chunks)
        return self._o._unpack(eo, xs.size(1))

    __call__ = forward


class B200Whisper(B200ASRModel):
    """Whisper (TransformerEncoder + TransformerDecoder [+ CTC]) on libwenet_b200.so; decode() keeps ASRModel.decode's
    signature (asr_model.py:267-343), `attention` being the mode Whisper supports (whisper.py:31)."""

    def __init__(self, configs: dict, state_dict: Dict[str, torch.Tensor], device=None, with_decoder: bool = True,
                 lang_table=WHISPER_LANGS, precise: bool = False):
        """precise=True: the parity mode of the ENCODER (bf16x3 GEMMs over [hi | lo | hi] activations, fp32 q / k / v and
        attention; encoder_out within 1e-3 of the fp32 reference, tests/test_whisper_gpu.py); the decoder stays bf16 as in
        the Conformer path."""
        super().__init__(configs, state_dict, device=device, with_decoder=with_decoder, precise=precise)
        assert self.spec.arch == 1, "B200Whisper needs a whisper-style configuration (encoder: transformer / conv1d2)"
        self.special_tokens = dict(self.spec.special_tokens)
        self.sos = self.special_tokens["sot"]
        self.eos = self.special_tokens["eot"]
        self.default_decode_method = "attention"
        self.decode_maxlen = self.spec.dec_max_len
        self.encoder = B200TransformerEncoder(self)
        self.ctc = B200CTC(self)
        self.lang_table = tuple(lang_table)
        # optional cap on the hypothesis length of decode mode "attention" (None: the reference's bound, encoder frames
        # + 1); an extension used by bench.py to pin the number of decoding steps of a synthetic-weight run
        self.max_decode_len = None
        self._has_ctc = "ctc.ctc_lo.weight" in state_dict

    @property
    def is_multilingual(self):
        return self.vocab_size >= 51865

    @property
    def num_languages(self):
        return self.vocab_size - 51765 - int(self.is_multilingual)

    def subsampling_rate(self) -> int:
        return 2

    def right_context(self) -> int:
        return 4

    def clone_shared(self):
        other = super().clone_shared()
        other.encoder = B200TransformerEncoder(other)
        return other

    @staticmethod
    def _out_frames(T: int) -> int:
        return (T - 1) // 2 + 1 if T > 0 else 0      # Conv1d(k3, s2, p1) output length

    def _encode(self, speech: torch.Tensor, speech_lengths: torch.Tensor, decoding_chunk_size: int = -1,
                num_decoding_left_chunks: int = -1) -> _EncOut:
        if not speech.is_cuda:
            raise _lib.WbError("speech must be a CUDA tensor (no CPU fallback)")
        x = speech.to(torch.float32).contiguous()
        B, T, D = x.shape
        assert D == self.spec.input_dim
        lens_host = _i32(speech_lengths.detach().cpu().numpy())
        assert int(lens_host.max()) <= T
        lib = self._lib
        rows = int(lib.wb_whisper_encoder_out_rows(B, ptr(lens_host), T))
        d = self.spec.d_model
        eo = _EncOut()
        eo.rows = rows
        eo.f32 = torch.empty(max(rows, 1), d, device=self.device, dtype=torch.float32)
        eo.bf16 = torch.empty(max(rows, 1), d * (3 if self.precise else 1), device=self.device, dtype=torch.bfloat16)
        eo.seq_start = torch.zeros(B, device=self.device, dtype=torch.int32)
        eo.seq_len = torch.zeros(B, device=self.device, dtype=torch.int32)
        tp = (lens_host // 2 if T % 2 == 0 else (lens_host + 1) // 2).astype(np.int32)     # subsampling.py:171
        eo.lens_host = tp
        eo.starts_host = np.concatenate([[0], np.cumsum(tp)[:-1]]).astype(np.int32)
        eo.max_len = int(tp.max()) if B else 0
        eo.dump = None
        if rows == 0:
            return eo
        wsb = lib.wb_whisper_encoder_workspace_bytes(self.dm.handle, B, ptr(lens_host), T)
        ws = self._workspace(wsb)
        with torch.cuda.device(self.device):
            check(lib.wb_whisper_encoder_forward(self.dm.handle, ptr(x), x.stride(0), ptr(lens_host), B, T, ptr(eo.f32),
                                                 ptr(eo.bf16), ptr(eo.seq_start), ptr(eo.seq_len), ptr(ws), ws.numel(),
                                                 cur_stream()), "wb_whisper_encoder_forward")
        return eo

    def _unpack(self, eo: _EncOut, T_in: int) -> Tuple[torch.Tensor, torch.Tensor]:
        B = eo.seq_start.numel()
        Tp = self._out_frames(T_in)
        d = self.spec.d_model
        out = torch.empty(B, Tp, d, device=self.device, dtype=torch.float32)
        if Tp > 0:
            check(self._lib.wb_unpack_rows(ptr(eo.f32), ptr(eo.seq_start), ptr(eo.seq_len), B, Tp, d, ptr(out), Tp,
                                           cur_stream()), "wb_unpack_rows")
        masks = (torch.arange(Tp, device=self.device).unsqueeze(0) < eo.seq_len.unsqueeze(1)).unsqueeze(1)
        return out, masks

    def _forward_chunk(self, *a, **k):
        raise NotImplementedError("streaming is not part of the Whisper path")

    def decode(self, methods: List[str], speech: torch.Tensor, speech_lengths: torch.Tensor, beam_size: int = 1,
               decoding_chunk_size: int = -1, num_decoding_left_chunks: int = -1, ctc_weight: float = 0.0,
               simulate_streaming: bool = False, reverse_weight: float = 0.0, context_graph=None, blank_id: int = 0,
               blank_penalty: float = 0.0, length_penalty: float = 0.0,
               infos: Dict[str, List[str]] = None) -> Dict[str, List[DecodeResult]]:
        assert speech.shape[0] == speech_lengths.shape[0]
        B = speech.shape[0]
        for m_ in methods:
            if m_ not in ("attention", "ctc_greedy_search", "ctc_prefix_beam_search"):
                raise NotImplementedError("decode mode %r is not available for Whisper (reverse_weight == 0, whisper.py:48)" % m_)
        if simulate_streaming:
            raise NotImplementedError("streaming is not part of the Whisper path")
        results = {}
        with torch.cuda.device(self.device):
            eo = self._encode(speech, speech_lengths)
            if "attention" in methods:
                if infos is None:                       # search.py:270-275
                    tasks, langs = ["transcribe"] * B, ["en"] * B
                else:
                    tasks, langs = infos["tasks"], infos["langs"]
                prefix = whisper_prefix(self.special_tokens, tasks, langs, self.lang_table)
                maxlen = self._out_frames(speech.size(1))
                if self.max_decode_len is not None:
                    maxlen = min(maxlen, int(self.max_decode_len))
                results["attention"] = self._attention_beam(eo, beam_size, length_penalty, prefix, self.eos, maxlen)
            rest = [m_ for m_ in methods if m_ != "attention"]
            if rest:
                if not self.spec_has_ctc():
                    raise _lib.WbError("this Whisper checkpoint carries no CTC head")
                need_beam = "ctc_prefix_beam_search" in rest
                logp, tv, ti = self._ctc(eo, beam_size if need_beam else 1, blank_id, blank_penalty, full=False)
                if "ctc_greedy_search" in rest:
                    results["ctc_greedy_search"] = self._greedy(eo, ti, blank_id)
                if need_beam:
                    bd = self._prefix_beam_launch(eo, tv, ti, beam_size, blank_id, None)
                    meta = self._beam_meta(bd)
                    fetch = self._beam_fetch_async(bd, meta)
                    results["ctc_prefix_beam_search"] = self._beam_results(bd, meta, fetch)
        return results

    def spec_has_ctc(self) -> bool:
        return bool(self.__dict__.get("_has_ctc", False))
