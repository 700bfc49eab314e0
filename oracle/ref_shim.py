"""Loader for the REAL reference projector (test infrastructure only).

Only usable where ``/root/reference`` is mounted (the authoring container).  It is
used by ``oracle/make_golden.py`` to produce the committed fixtures under
``tests/golden/`` and by the CPU tests that pin the oracle restatement against the
reference itself.  Nothing on the GPU box may rely on it: ``load_reference()``
returns ``None`` when the mount is absent.

Recipe: the reference's ``hicom/__init__.py`` drags in the whole
model zoo and ``projector.py:13`` imports ``TRANSFORMERS_CACHE`` (gone in
transformers 5.x), and ``hicom/mm_utils.py:10-15`` wants decord / moviepy /
imageio.  We register empty stand-ins for those, a bare ``hicom`` package object
whose ``__path__`` points at the mount, and load ``hicom/model/projector.py`` by path.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "hicom", "model", "projector.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_CACHED = None


def load_reference():
    """Return the reference ``hicom.model.projector`` module, or None if not mounted."""
    global _CACHED
    if _CACHED is not None:
        return _CACHED
    if not reference_available():
        return None
    import transformers

    if not hasattr(transformers, "TRANSFORMERS_CACHE"):
        transformers.TRANSFORMERS_CACHE = "/tmp/_hf_cache_unused"
    _stub("decord", VideoReader=object, cpu=lambda *a, **k: None)
    _stub("imageio")
    mp = _stub("moviepy")
    ed = _stub("moviepy.editor", VideoFileClip=object)
    mp.editor = ed
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            _stub("cv2")

    pkg = types.ModuleType("hicom")
    pkg.__path__ = [os.path.join(REF_ROOT, "hicom")]
    sys.modules.setdefault("hicom", pkg)
    mpkg = types.ModuleType("hicom.model")
    mpkg.__path__ = [os.path.join(REF_ROOT, "hicom", "model")]
    sys.modules.setdefault("hicom.model", mpkg)

    path = os.path.join(REF_ROOT, "hicom", "model", "projector.py")
    spec = importlib.util.spec_from_file_location("hicom.model.projector", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hicom.model.projector"] = mod
    spec.loader.exec_module(mod)
    _CACHED = mod
    return mod


class BagConfig:
    """Plain attribute bag standing in for the HF config the reference reads."""

    def __init__(self, **kw):
        self.mm_vision_tower = "google/siglip-so400m-patch14-384"
        self.mm_hidden_size = 1152
        self.hidden_size = 896
        self.mm_projector_type = "local43_global32"
        for k, v in kw.items():
            setattr(self, k, v)
