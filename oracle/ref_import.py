"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference hot path.

Loads /root/reference/speech_decoding/{models.py,utils/loss.py} as they lie on
disk (nothing is copied into this repo), with three stub modules for
third-party imports that are absent in this image (termcolor, mne, mne_bids)
and a synthetic sensor layout substituted for `ch_locations_2d`
(reference: speech_decoding/utils/layout.py:6-43 needs mne + the dataset).

/root/reference exists only in the build container; oracle/install_reference.py pip-installs the same unmodified
package into baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot).  This module is used to
  * validate oracle/restate.py (tests, CPU, when the reference is present),
  * generate tests/golden/*.npz (oracle/gen_golden.py), and
  * give bench.py its reference arm / cpu_baseline / stock-eager baseline (the reference's own modules, timed --
    never on the product path; bench.py falls back to the oracle port when neither location exists).
Nothing in the product path, the -m gpu tests or smoke() imports it.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("SD_REFERENCE_ROOT", "/root/reference"),
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


def _find():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "speech_decoding", "models.py")):
            return c
    return _CANDIDATES[0]


REFERENCE_ROOT = _find()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "speech_decoding", "models.py"))


def where() -> str:
    """'reference' (the mounted tree) or '_ref' (the pip-installed copy under baseline/_ref)."""
    return "_ref" if REFERENCE_ROOT.endswith(os.path.join("baseline", "_ref")) else "reference"


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
