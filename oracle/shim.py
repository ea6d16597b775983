"""Oracle shim: make the UNMODIFIED reference (``/root/reference``) importable in the
build container so it can (a) pin the CPU restatement in ``oracle/wenet_oracle.py`` and
(b) generate the golden fixtures under ``tests/golden/`` (``oracle/make_goldens.py``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``wenet_b200/`` may import this module.
``/root/reference`` does not exist on the GPU box, so nothing that runs there may need it:
``have_reference()`` is the guard.

The five shims are the ones SURVEY.md section 8(c) found empirically:
  1. ``librosa``  missing  (wenet/dataset/processor.py:23, only used by whisper log-mel)
  2. ``langid``   missing  (wenet/dataset/processor.py:28,35,110-112)
  3. ``whisper.tokenizer.LANGUAGES`` missing (wenet/utils/common.py:24)
  4. ``torch.nn.modules.conv.Union/Optional`` gone in torch 2.11
     (wenet/models/squeezeformer/conv2d.py:17, pulled in by wenet/utils/class_utils.py:14)
  5. ``torchaudio.load`` needs torchcodec (wenet/dataset/processor.py:141-148)
"""
import os
import sys
import types
import typing
import wave

def _find_reference():
    """WENET_REFERENCE_ROOT, else the read-only tree of the build container, else the pip --target install of the
    same unmodified sources under baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)."""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for c in (os.environ.get("WENET_REFERENCE_ROOT"), "/root/reference", os.path.join(repo, "baseline", "_ref")):
        if c and os.path.isdir(os.path.join(c, "wenet")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "wenet"))


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Idempotent.  Must run before the first ``import wenet``."""
    global _installed
    if _installed:
        return
    if not have_reference():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    import torch
    import torch.nn.modules.conv as _c
    import torchaudio

    for n in ("Union", "Optional"):
        if not hasattr(_c, n):
            setattr(_c, n, getattr(typing, n))
    if "whisper" not in sys.modules:
        w = _stub("whisper")
        w.tokenizer = _stub("whisper.tokenizer",
                            LANGUAGES={"en": "english", "zh": "chinese"})
    if "librosa" not in sys.modules:
        _stub("librosa")

    class _LID:
        @classmethod
        def from_modelstring(cls, *a, **k):
            return cls()

        def set_languages(self, langs):
            pass

        def classify(self, txt):
            return ("zh", 1.0)

    if "langid" not in sys.modules:
        lid = _stub("langid")
        lid.langid = _stub("langid.langid", LanguageIdentifier=_LID, model="")

    def _load(f, **kw):
        wf = wave.open(f, "rb")
        a = np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16)
        return (torch.from_numpy(a.astype(np.float32) / 32768.0).unsqueeze(0),
                wf.getframerate())

    torchaudio.load = _load
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def init_reference_model(configs: dict):
    """Build the reference model exactly as wenet/utils/init_model.py:184-217 does
    (no checkpoint, no jit)."""
    install()
    from wenet.utils.init_model import init_model
    args = types.SimpleNamespace(checkpoint=None, jit=False, enc_init=None,
                                 enc_init_mods=None, freeze_modules=None,
                                 lora_ckpt_path=None, use_lora=False,
                                 only_optimize_lora=False)
    model, configs = init_model(args, configs)
    model.eval()
    return model
