"""On-GPU ingest and length-bucketed batching for decoding.

The reference feeds `recognize.py` through `wenet/dataset/dataset.py:26-155` + `processor.py`: per utterance
`decode_wav` (:125-196) -> `compute_fbank` on the CPU (~22 ms per 10 s utterance) -> `sort` / `static_batch` /
`padding` (:479-577), all inside DataLoader workers.  With the encoder at > 100 000 x real time that host pipeline
is what a real run waits for, so this module replaces it for inference:

    wav files --(reader thread: stdlib `wave`, int16, into PINNED host buffers)--> one H2D copy per batch
              --> fused fbank kernel on the device --> B200ASRModel.decode (packed rows, no padding work)

* `plan_batches`   length-bucketed batching: utterances sorted by length (as `processor.sort`, :479-503, but over the
                   whole list), greedily packed under a budget of padded audio seconds and a maximum batch size - the
                   dynamic-batch rule of `processor.dynamic_batch` (:545-577) applied to samples instead of frames.
* `WavBatcher`     the reader thread: decodes and pins batch k + 1 while batch k is on the GPU.
* `transcribe_files` the whole pipeline; returns {key: DecodeResult} in input order.

Only 16-bit PCM RIFF files at the model's sample rate are accepted (what `recognize.py` is fed after the recipes'
data preparation); anything else raises - resampling / other codecs stay with the reference's `torchaudio` path.
"""
import queue
import threading
import wave
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .fbank import FbankExtractor


def read_wav_int16(path: str, sample_rate: int = 16000) -> np.ndarray:
    """16-bit PCM RIFF -> int16 samples of channel 0 (processor.decode_wav keeps the first channel, :141-150)."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError("%s: %d-bit samples (only 16-bit PCM is accepted)" % (path, 8 * w.getsampwidth()))
        if w.getframerate() != sample_rate:
            raise ValueError("%s: %d Hz (the model expects %d Hz; resample with the reference pipeline)"
                             % (path, w.getframerate(), sample_rate))
        ch = w.getnchannels()
        a = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    return np.ascontiguousarray(a[::ch] if ch > 1 else a)


def wav_num_samples(path: str) -> int:
    with wave.open(path, "rb") as w:
        return w.getnframes()


def plan_batches(num_samples: Sequence[int], max_batch_seconds: float = 1920.0, max_batch_size: int = 64,
                 sample_rate: int = 16000) -> List[List[int]]:
    """Indices grouped into batches: longest first, a batch closes when adding the next utterance would push
    (batch size) x (longest utterance in the batch) past the budget or the size limit.  Every index appears once."""
    order = sorted(range(len(num_samples)), key=lambda i: (-int(num_samples[i]), i))
    budget = int(max_batch_seconds * sample_rate)
    batches, cur, cur_max = [], [], 0
    for i in order:
        n = int(num_samples[i])
        longest = max(cur_max, n)
        if cur and (len(cur) + 1 > max_batch_size or (len(cur) + 1) * longest > budget):
            batches.append(cur)
            cur, longest = [], n
        cur.append(i)
        cur_max = longest
    if cur:
        batches.append(cur)
    return batches


class WavBatcher:
    """Iterates (indices, pinned int16 PCM (B, n_max), num_samples (B,)) over `plan_batches`' batches; a background
    thread reads and pins the next batches while the caller works on the current one."""

    def __init__(self, paths: Sequence[str], batches: List[List[int]], sample_rate: int = 16000, prefetch: int = 2):
        self.paths, self.batches, self.sr = list(paths), batches, sample_rate
        self.q: "queue.Queue" = queue.Queue(maxsize=max(1, prefetch))
        self.err = None
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        try:
            for idx in self.batches:
                rows = [read_wav_int16(self.paths[i], self.sr) for i in idx]
                n_max = (max(r.shape[0] for r in rows) + 7) // 8 * 8
                buf = torch.zeros(len(rows), n_max, dtype=torch.int16).pin_memory()
                for b, r in enumerate(rows):
                    buf[b, :r.shape[0]] = torch.from_numpy(r)
                ns = torch.tensor([r.shape[0] for r in rows], dtype=torch.int32)
                self.q.put((idx, buf, ns))
        except Exception as e:  # noqa: BLE001 - re-raised in the consumer
            self.err = e
        finally:
            self.q.put(None)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                if self.err is not None:
                    raise self.err
                return
            yield item


def transcribe_files(model, paths: Sequence[str], keys: Sequence[str] = None, methods: Iterable[str] = ("attention_rescoring",),
                     beam_size: int = 10, ctc_weight: float = 0.5, reverse_weight: float = 0.0,
                     max_batch_seconds: float = 1920.0, max_batch_size: int = 64, decoding_chunk_size: int = -1,
                     num_decoding_left_chunks: int = -1, num_mel_bins: int = 80) -> Dict[str, Dict[str, object]]:
    """Decode wav files with a B200ASRModel: {method: {key: DecodeResult}}.  The arguments after `keys` are those of
    ASRModel.decode (asr_model.py:267-283) plus the two batching limits."""
    if not torch.cuda.is_available():
        raise _lib.WbError("transcribe_files needs a CUDA device (no CPU fallback)")
    paths = list(paths)
    keys = list(keys) if keys is not None else paths
    assert len(keys) == len(paths)
    methods = list(methods)
    fb = FbankExtractor(num_mel_bins)
    ns_all = [wav_num_samples(p) for p in paths]
    too_short = [k for k, n in zip(keys, ns_all) if fb.num_frames(n) < 7]
    if too_short:
        raise ValueError("utterances shorter than 7 frames give no encoder output: %s" % too_short[:5])
    batches = plan_batches(ns_all, max_batch_seconds, max_batch_size)
    out: Dict[str, Dict[str, object]] = {m: {} for m in methods}
    dev = model.device
    with torch.cuda.device(dev):
        for idx, pcm_host, ns in WavBatcher(paths, batches):
            pcm = pcm_host.to(dev, non_blocking=True)
            nsd = ns.to(dev, non_blocking=True)
            feats = fb(pcm, nsd)
            flens = torch.tensor([fb.num_frames(int(n)) for n in ns], dtype=torch.int64, device=dev)
            res = model.decode(methods, feats[:, :int(flens.max())], flens, beam_size=beam_size, ctc_weight=ctc_weight,
                               reverse_weight=reverse_weight, decoding_chunk_size=decoding_chunk_size,
                               num_decoding_left_chunks=num_decoding_left_chunks)
            for m in methods:
                for b, i in enumerate(idx):
                    out[m][keys[i]] = res[m][b]
    return {m: {k: out[m][k] for k in keys} for m in methods}
