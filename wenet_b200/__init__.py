"""wenet_b200 — B200-native (sm_100a) Conformer ASR inference hot path behind WeNet's Python API.

Public surface (mirrors the reference's names):
    install()                      rebinding of WeNet's registries / compute_fbank
    B200ASRModel.from_reference()  wrap a loaded reference model
    compute_fbank                  drop-in for wenet.dataset.processor.compute_fbank
The compute path is libwenet_b200.so (hand-written CUDA, C ABI in include/wenet_b200.h); there is
no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def __getattr__(name):
    import importlib
    lazy = {
        "compute_fbank": ("fbank", "compute_fbank"),
        "FbankExtractor": ("fbank", "FbankExtractor"),
        "B200ASRModel": ("asr_model", "B200ASRModel"),
        "B200ConformerEncoder": ("asr_model", "B200ConformerEncoder"),
        "StreamingSession": ("asr_model", "StreamingSession"),
        "DecodeResult": ("search", "DecodeResult"),
        "install": ("plugin", "install"),
        "transcribe_files": ("ingest", "transcribe_files"),
        "DeviceModel": ("weights", "DeviceModel"),
        "ModelSpec": ("weights", "ModelSpec"),
    }
    if name in lazy:
        mod, attr = lazy[name]
        return getattr(importlib.import_module("." + mod, __name__), attr)
    raise AttributeError(name)
