"""Host side of the fused fbank kernel (csrc/fbank.cu).

`compute_fbank(sample, ...)` keeps the signature and semantics of the reference's
wenet/dataset/processor.py:226-256 (which calls torchaudio.compliance.kaldi.fbank with dither 0 in
decoding, energy_floor 0, povey window) so `wenet_b200.install()` can rebind it; `FbankExtractor`
is the batched device API the benchmark uses (PCM already in HBM -> (B, T, 80) features).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr


def _povey_window(n: int) -> torch.Tensor:
    # torchaudio/compliance/kaldi.py:99-100
    return torch.hann_window(n, periodic=False, dtype=torch.float32).pow(0.85)


def _mel_banks(num_bins: int, padded: int, sample_freq: float, low_freq: float = 20.0,
               high_freq: float = 0.0) -> torch.Tensor:
    """torchaudio/compliance/kaldi.py:436-511 (vtln_warp = 1) with the zero column of :627 appended;
    same float32 torch arithmetic so the filter weights are bit-identical."""
    num_fft_bins = padded / 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return torch.nn.functional.pad(bins, (0, 1)).to(torch.float32).contiguous()


class FbankExtractor:
    """One fused kernel: frames -> DC removal -> pre-emphasis -> povey -> rFFT -> mel -> log."""

    def __init__(self, num_mel_bins=80, frame_length=25, frame_shift=10, sample_rate=16000,
                 window_type="povey", preemph=0.97):
        if window_type != "povey":
            raise NotImplementedError("only the povey window (wenet's default) is implemented")
        self.num_mel = int(num_mel_bins)
        self.frame_len = int(sample_rate * frame_length * 0.001)
        self.frame_shift = int(sample_rate * frame_shift * 0.001)
        self.nfft = 1 << (self.frame_len - 1).bit_length()
        win = _povey_window(self.frame_len).numpy().astype(np.float32)
        mel = _mel_banks(self.num_mel, self.nfft, float(sample_rate)).numpy().astype(np.float32)
        self._h = C.c_void_p()
        check(_lib.load().wb_fbank_create(C.byref(self._h), self.num_mel, self.frame_len, self.frame_shift,
                                          float(preemph), ptr(win), ptr(mel)), "wb_fbank_create")

    def __del__(self):
        try:
            if self._h:
                _lib.load().wb_fbank_destroy(self._h)
        except Exception:
            pass

    def num_frames(self, n: int) -> int:
        return 1 + (n - self.frame_len) // self.frame_shift if n >= self.frame_len else 0

    def __call__(self, pcm: torch.Tensor, num_samples: torch.Tensor, scale: float = None,
                 max_frames: int = None) -> torch.Tensor:
        """pcm (B, N) float32 (in [-1,1), scale defaults to 32768) or int16 (scale 1) on the GPU;
        num_samples (B,) int32 on the GPU.  Returns (B, max_frames, num_mel) float32."""
        if not pcm.is_cuda:
            raise _lib.WbError("FbankExtractor needs CUDA tensors (no CPU fallback)")
        assert pcm.dim() == 2 and pcm.stride(1) == 1
        is_i16 = pcm.dtype == torch.int16
        if not is_i16 and pcm.dtype != torch.float32:
            raise TypeError("pcm must be float32 or int16")
        if scale is None:
            scale = 1.0 if is_i16 else 32768.0
        B, N = pcm.shape
        if max_frames is None:
            max_frames = self.num_frames(N)
        out = torch.empty(B, max_frames, self.num_mel, device=pcm.device, dtype=torch.float32)
        ns = num_samples.to(device=pcm.device, dtype=torch.int32)
        check(_lib.load().wb_fbank_forward(self._h, ptr(pcm), int(is_i16), pcm.stride(0), ptr(ns), B, float(scale),
                                           ptr(out), max_frames, max_frames, cur_stream()), "wb_fbank_forward")
        return out


_extractors = {}


def _get_extractor(num_mel_bins, frame_length, frame_shift, sample_rate, window_type):
    key = (num_mel_bins, frame_length, frame_shift, sample_rate, window_type)
    if key not in _extractors:
        _extractors[key] = FbankExtractor(num_mel_bins, frame_length, frame_shift, sample_rate, window_type)
    return _extractors[key]


def compute_fbank(sample, num_mel_bins=23, frame_length=25, frame_shift=10, dither=0.0, window_type="povey"):
    """Drop-in for wenet.dataset.processor.compute_fbank (processor.py:226-256).

    sample: {key, wav (1, n) float32 in [-1, 1), sample_rate, ...} -> adds 'feat' (m, num_mel_bins)
    float32 (on the CPU, like the reference, so the reference's padding/collate code keeps working).
    dither must be 0 (recognize.py:225-226 forces it for decoding)."""
    assert "sample_rate" in sample and "wav" in sample and "key" in sample
    if dither != 0.0:
        raise NotImplementedError("dither is a training-time augmentation; decoding uses dither=0")
    wav = sample["wav"]
    dev = torch.device("cuda", torch.cuda.current_device())
    ex = _get_extractor(num_mel_bins, frame_length, frame_shift, int(sample["sample_rate"]), window_type)
    x = wav[0:1].to(device=dev, dtype=torch.float32).contiguous()
    n = torch.tensor([x.shape[1]], dtype=torch.int32, device=dev)
    sample["feat"] = ex(x, n)[0].cpu()
    return sample
