"""TEST / BENCH INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference package into baseline/_ref/.

    python oracle/install_reference.py        (also run by __graft_entry__.build() when /root/reference is mounted)

baseline/_ref/ is git-ignored (no reference source ever enters this repository's history) but NOT gpurun-ignored, so
the installed copy travels to the GPU box with the snapshot; there `bench.py --impl reference`, the `cpu_baseline`
leg and the stock-eager GPU baseline run the reference's own modules (through oracle/ref_import.py) instead of the
oracle port.

Recipe (the contract's one offline install):
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
The reference ships neither setup.py nor pyproject.toml (it is run from its checkout: README.md), so that command
stops with "does not appear to be a Python project".  The fallback copies the tree to a scratch directory, adds a
five-line setup.py there (packaging metadata only -- no module is touched) and installs that copy with --no-deps
(its requirements pin torch==1.12.1+cu113 and friends, which are neither installable nor wanted here).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SD_REFERENCE_ROOT", "/root/reference")
TARGET = os.path.join(ROOT, "baseline", "_ref")

SETUP_PY = """from setuptools import setup, find_namespace_packages
setup(name="speech_decoding_reference", version="0", packages=find_namespace_packages(include=["speech_decoding", "speech_decoding.*"]))
"""


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "speech_decoding", "models.py"))


def _pip(src, extra=()):
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
           "--target", TARGET, "--upgrade", "-q", *extra, src]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def install(verbose=True) -> str:
    """Returns a one-line outcome string (recorded in DESIGN.md)."""
    if not os.path.isfile(os.path.join(REF, "speech_decoding", "models.py")):
        return "reference tree not mounted at %s: kept what is in baseline/_ref (%s)" % (REF, "present" if installed() else "absent")
    os.makedirs(TARGET, exist_ok=True)
    r = _pip(REF)
    how = "direct"
    if r.returncode != 0:
        tmp = tempfile.mkdtemp(prefix="sd_ref_")
        try:
            dst = os.path.join(tmp, "reference")
            shutil.copytree(REF, dst, ignore=shutil.ignore_patterns("data", "assets", ".git"))
            with open(os.path.join(dst, "setup.py"), "w") as f:
                f.write(SETUP_PY)
            r = _pip(dst, ("--no-deps",))
            how = "from a scratch copy with a metadata-only setup.py, --no-deps"
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if r.returncode != 0 or not installed():
        return "pip install of the reference FAILED: " + " | ".join((r.stdout or "").strip().splitlines()[-2:])
    # the installed modules must be byte-identical to the mounted reference
    for rel in ("models.py", os.path.join("utils", "loss.py"), os.path.join("utils", "preproc_utils.py")):
        a = open(os.path.join(REF, "speech_decoding", rel), "rb").read()
        b = open(os.path.join(TARGET, "speech_decoding", rel), "rb").read()
        assert a == b, rel + " differs from the reference"
    msg = "reference installed into baseline/_ref (%s); modules byte-identical to %s" % (how, REF)
    if verbose:
        print(msg)
    return msg


if __name__ == "__main__":
    print(install(verbose=False))
