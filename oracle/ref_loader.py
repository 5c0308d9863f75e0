"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE ONLY.  The reference pulls in ``tensorflow`` and ``gensim``
at import time (libreco/tfops/version.py:1-14, libreco/bases/gensim_base.py:5);
neither is installed, so two stub modules are registered in ``sys.modules``
first.  The numpy / torch half of the library then imports and runs unchanged
(SURVEY.md §0.3).  ``/root/reference`` does not exist on the GPU box: there the byte-identical
copy staged by ``oracle/make_ref.py`` under ``oracle/_ref`` (git-ignored, shipped with the snapshot)
is imported instead.  Callers must guard with :func:`reference_available`.
"""
import os
import sys
import types
from importlib.machinery import ModuleSpec
from unittest.mock import MagicMock

_MOUNTED = os.environ.get("B200RECO_REFERENCE", "/root/reference")
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = _MOUNTED if os.path.isdir(os.path.join(_MOUNTED, "libreco")) else _STAGED


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "libreco"))


def reference_kind() -> str:
    """"mounted" (/root/reference), "staged" (oracle/_ref) or "absent"."""
    if not reference_available():
        return "absent"
    return "mounted" if REFERENCE_ROOT == _MOUNTED else "staged"


def sample_data_path(name="sample_movielens_rating.dat") -> str:
    return os.path.join(REFERENCE_ROOT, "examples", "sample_data", name)


def _install_stubs() -> None:
    if "tensorflow" not in sys.modules:
        tf = types.ModuleType("tensorflow")
        tf.__version__ = "2.12.0"
        tf.__spec__ = ModuleSpec("tensorflow", None)   # torch._dynamo probes find_spec("tensorflow")

        class _V1(MagicMock):
            __version__ = "2.12.0"  # read as TF_VERSION by tfops/version.py:6

        tf.compat = types.SimpleNamespace(v1=_V1())
        compat = types.ModuleType("tensorflow.compat")
        compat.__spec__ = ModuleSpec("tensorflow.compat", None)
        compat.v1 = tf.compat.v1
        sys.modules["tensorflow"] = tf
        sys.modules["tensorflow.compat"] = compat
        sys.modules["tensorflow.compat.v1"] = tf.compat.v1
    if "gensim" not in sys.modules:
        gensim = types.ModuleType("gensim")
        models = types.ModuleType("gensim.models")
        gensim.__spec__ = ModuleSpec("gensim", None)
        models.__spec__ = ModuleSpec("gensim.models", None)

        class Word2Vec:  # never instantiated on the paths we run
            pass

        models.Word2Vec = Word2Vec
        gensim.models = models
        sys.modules["gensim"] = gensim
        sys.modules["gensim.models"] = models


def load_reference():
    """Return the imported ``libreco`` package of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    import libreco  # noqa: F401

    return libreco
