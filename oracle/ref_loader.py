"""Import the UNMODIFIED reference (``/root/reference``) under inert stubs.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only where
``/root/reference`` exists (the build container); the GPU box never has it, so
nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may call this at run time.

The reference needs ten third-party packages that are absent here
(pyroomacoustics, soundfile, librosa, matplotlib, noisereduce, mir_eval,
asteroid, speechbrain, opuslib, seaborn).  All are used for I/O, plotting,
metrics or training, except ``pyroomacoustics.transform.stft.analysis`` which
is on the hot path and is replaced by the A1 restatement in ``pra_stft.py``.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ASW_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "soundfile", "librosa", "librosa.core", "librosa.display", "matplotlib",
    "matplotlib.pyplot", "matplotlib.patches", "noisereduce", "mir_eval",
    "mir_eval.separation", "asteroid", "asteroid.losses",
    "asteroid.losses.sdr", "asteroid.metrics", "speechbrain",
    "speechbrain.lobes", "speechbrain.lobes.models",
    "speechbrain.lobes.models.transformer",
    "speechbrain.lobes.models.transformer.Conformer",
    "speechbrain.lobes.models.transformer.Transformer",
    "speechbrain.nnet", "speechbrain.nnet.schedulers", "opuslib", "seaborn",
    "pyroomacoustics", "pyroomacoustics.transform",
    "pyroomacoustics.transform.stft",
]


class _Inert(types.ModuleType):
    """Module whose unknown attributes are inert callables/classes."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)

        class _Dummy:  # usable as a base class, a callable, or a namespace
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                return None

            def __getattr__(self, n):
                if n.startswith("__") and n.endswith("__"):
                    raise AttributeError(n)
                return _Dummy()

        _Dummy.__name__ = name
        return _Dummy


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sep"))


def load():
    """Return a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    # real heavy deps first so the stubs cannot shadow anything they import
    import numpy  # noqa: F401
    import scipy.ndimage  # noqa: F401
    import torch  # noqa: F401
    try:
        import torchaudio  # noqa: F401
    except Exception:  # pragma: no cover
        sys.modules.setdefault("torchaudio", _Inert("torchaudio"))

    from . import pra_stft

    for name in _STUBS:
        if name in sys.modules and not isinstance(sys.modules[name], _Inert):
            continue  # a real install wins
        try:
            if importlib.util.find_spec(name.split(".")[0]) is not None and name.split(".")[0] not in (
                    "pyroomacoustics",):
                # real package present: do not stub it
                continue
        except (ImportError, ValueError):
            pass
        mod = _Inert(name)
        mod.__path__ = []  # behave like a package
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)
    # the one arithmetic dependency: A1
    sys.modules["pyroomacoustics.transform.stft"].analysis = pra_stft.analysis

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    ns = types.SimpleNamespace()
    ns.constants = importlib.import_module("sep.helpers.constants")
    ns.Patch_3D = importlib.import_module("sep.Traditional_SP.Patch_3D")
    ns.SRP_Prunning = importlib.import_module("sep.Traditional_SP.SRP_Prunning")
    ns.local_utils_3d = importlib.import_module("sep.helpers.local_utils_3d")
    ns.Mic_Array = importlib.import_module("sep.Mic_Array")
    ns.spot_network = importlib.import_module("sep.training.SpeakerLocalization.network")
    try:
        ns.joint_network = importlib.import_module("sep.training.JointModel.network")
    except Exception as e:  # pragma: no cover - depends on stubs being enough
        ns.joint_network = None
        ns.joint_network_error = repr(e)
    return ns
