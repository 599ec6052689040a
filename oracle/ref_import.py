"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference class in the build container.

``/root/reference`` exists only in the build container, never on the GPU box, so this
module is used solely by ``oracle/make_golden.py`` (fixture generation) and by CPU tests
that skip when the tree is absent.  Recipe from SURVEY.md §8c: the reference's
``tal.asr.models`` imports unrelated packages that are not installed here; they are
replaced by empty stand-ins, and the historical package name ``wildspeech`` is aliased
to ``tal`` (tal/asr/models.py:12).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TALFE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "tal", "asr", "models.py")):
        return False
    try:
        import torchaudio  # noqa: F401
    except Exception:
        return False
    return True


def _stand_in(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load_reference_models():
    """Returns the reference's ``tal.asr.models`` module (LogMelSpec, ASRModel, SDModel...)."""
    import torch

    class _Unused:                      # instantiated at import time by tal/asr/data/util.py:9
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, item):
            raise RuntimeError("stand-in for a package the log-mel path never touches")

    _stand_in("librosa")
    _stand_in("librosa.core", get_duration=None)
    _stand_in("librosa.effects", time_stretch=None)
    _stand_in("mutagen", File=None)
    _stand_in("nltk")
    _stand_in("nltk.tokenize", TweetTokenizer=_Unused, word_tokenize=None)
    _stand_in("rezero")
    _stand_in("rezero.transformer", RZTXDecoderLayer=torch.nn.Module)
    _stand_in("fairseq")
    _stand_in("fairseq.models")
    _stand_in("fairseq.models.wav2vec", Wav2VecModel=_Unused)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import tal
    import tal.modules
    sys.modules.setdefault("wildspeech", tal)
    sys.modules.setdefault("wildspeech.modules", tal.modules)
    import tal.asr.models as ref_models
    return ref_models


def reference_logmel(double: bool = False, **ctor):
    """An instance of the reference's own LogMelSpec (fp32), or its float64 twin; ``ctor``: sr / n_mels / eps as in
    tal/asr/models.py:22."""
    mod = load_reference_models().LogMelSpec(**ctor)
    return mod.double() if double else mod
