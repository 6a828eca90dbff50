"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference hot path.

Loads /root/reference/speech_decoding/{models.py,utils/loss.py} as they lie on
disk (nothing is copied into this repo), with three stub modules for
third-party imports that are absent in this image (termcolor, mne, mne_bids)
and a synthetic sensor layout substituted for `ch_locations_2d`
(reference: speech_decoding/utils/layout.py:6-43 needs mne + the dataset).

/root/reference exists only in the build container.  This module is used to
  * validate oracle/restate.py (tests, CPU, when the reference is present), and
  * generate tests/golden/*.npz (oracle/gen_golden.py).
Nothing in the product path, the -m gpu tests, smoke() or bench.py imports it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "speech_decoding", "models.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def load(layout_fn):
    """Return (models_module, loss_module) of the unmodified reference.

    layout_fn(args) -> float32 tensor (C, 2): stands in for ch_locations_2d
    (bound by name at import, models.py:11, called at models.py:29).
    """
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("termcolor", cprint=lambda *a, **k: None)
    _stub("mne")
    _stub("mne_bids")
    # The drop-in package in this repo has the same import name; make sure the
    # reference's copy wins inside this (test-only) process.
    for k in [k for k in sys.modules if k == "speech_decoding" or k.startswith("speech_decoding.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        models = importlib.import_module("speech_decoding.models")
        loss = importlib.import_module("speech_decoding.utils.loss")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    # keep private handles, then free the import name for the drop-in again
    ref_models, ref_loss = models, loss
    for k in [k for k in sys.modules if k == "speech_decoding" or k.startswith("speech_decoding.")]:
        del sys.modules[k]
    ref_models.ch_locations_2d = layout_fn
    return ref_models, ref_loss


def load_preproc_utils():
    """The unmodified speech_decoding/utils/preproc_utils.py (baseline_correction_single :128-142, scaleAndClamp
    :69-90 -- what Gwilliams2022Collator.forward calls, dataclass/gwilliams2022.py:653-661).  Its imports of
    termcolor and omegaconf (absent in this image, unused by those two functions) are stubbed."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("termcolor", cprint=lambda *a, **k: None)
    _stub("omegaconf", open_dict=lambda *a, **k: None)
    import importlib.util
    path = os.path.join(REFERENCE_ROOT, "speech_decoding", "utils", "preproc_utils.py")
    spec = importlib.util.spec_from_file_location("_ref_preproc_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
