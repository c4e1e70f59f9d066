"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product (cocos_b200/).

Stages the UNMODIFIED reference (michaelnowotny/cocos) from /root/reference into
``oracle/_ref/`` so that it travels to the GPU box with the snapshot (``oracle/_ref/``
is git-ignored -- nothing of the reference enters the history -- but not
gpurun-ignored, like the built ``.so`` files).  ``bench.py --impl reference`` and the
``cpu_baseline`` leg then time the reference's own ``simulate_and_compute_option_price``
(examples/stochastic_volatility/heston_utilities.py:122-164) on its NumPy bundle,
imported from ``oracle/_ref`` under the stubs of ``oracle/reference_runner.py``.

What is staged: exactly the reference files that importing its Heston entry point on
the NumPy bundle loads (24 pure-Python files: the ``cocos`` package's import closure and
``examples/stochastic_volatility/heston_utilities.py`` -- the examples live outside the
installable package, which is why ``pip install --target`` alone would not do); the
closure is determined at staging time by importing the reference in a subprocess.
Tests, notebooks, images and unrelated modules are not staged.  A manifest with the
SHA-256 of every staged file is written next to them; ``verify()`` recomputes it, so a
staged copy that was edited is detected.

    python oracle/build_ref.py            # stage (idempotent)
    python oracle/build_ref.py --verify   # check the staged copy against its manifest
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE_ROOT = "/root/reference"
STAGED_ROOT = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(STAGED_ROOT, "MANIFEST.json")


def _sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for chunk in iter(lambda: fh.read(1 << 16), b""):
            h.update(chunk)
    return h.hexdigest()


def _python_files(root):
    """Relative paths of the reference files its NumPy Heston path imports (found by importing it in a subprocess
    under the stubs of reference_runner and listing the modules that were loaded from `root`)."""
    import subprocess
    code = ("import os, sys\n"
            f"sys.path.insert(0, {os.path.dirname(HERE)!r})\n"
            "from oracle import reference_runner as rr\n"
            f"rr.REFERENCE_ROOT = {root!r}\n"
            "rr.load_reference()\n"
            f"root = {root!r} + os.sep\n"
            "files = {os.path.relpath(m.__file__, root) for m in list(sys.modules.values())\n"
            "         if getattr(m, '__file__', None) and m.__file__.startswith(root)}\n"
            "print('\\n'.join(sorted(files)))\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    return [line.strip() for line in out.splitlines() if line.strip().endswith(".py")]


def source_available():
    return os.path.isdir(os.path.join(SOURCE_ROOT, "cocos"))


def staged_available():
    return os.path.isfile(MANIFEST) and os.path.isdir(os.path.join(STAGED_ROOT, "cocos"))


def stage(verbose=True):
    """Copy the reference's Python files, byte for byte, into oracle/_ref/ and write the manifest."""
    if not source_available():
        if verbose:
            print(f"build_ref: {SOURCE_ROOT} is not mounted; keeping whatever is staged "
                  f"({'present' if staged_available() else 'nothing'})")
        return staged_available()
    manifest = {}
    if os.path.isdir(STAGED_ROOT):
        shutil.rmtree(STAGED_ROOT)                      # never keep files of an earlier, wider staging
    for rel in _python_files(SOURCE_ROOT):
        src, dst = os.path.join(SOURCE_ROOT, rel), os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        digest = _sha256(src)
        if not (os.path.exists(dst) and _sha256(dst) == digest):
            shutil.copyfile(src, dst)
        manifest[rel] = digest
    with open(MANIFEST, "w") as fh:
        json.dump({"source": SOURCE_ROOT, "files": manifest}, fh, indent=1, sort_keys=True)
    with open(os.path.join(STAGED_ROOT, "README.txt"), "w") as fh:
        fh.write("Byte-for-byte staged copy of the reference (michaelnowotny/cocos: the cocos package and the examples\n"
                 "tree), written by oracle/build_ref.py from /root/reference for ONE purpose: bench.py's CPU arm imports\n"
                 "and times the reference's own, unmodified NumPy path on the GPU box, where /root/reference does not\n"
                 "exist.  This directory is git-ignored: it is not part of the repository and nothing of it is product\n"
                 "code.  MANIFEST.json holds the SHA-256 of every file; `python oracle/build_ref.py --verify` checks it.\n")
    if verbose:
        print(f"build_ref: staged {len(manifest)} reference files under {STAGED_ROOT}")
    return True


def verify():
    """True when every staged file still has the digest recorded at staging time (i.e. is unmodified)."""
    if not staged_available():
        return False
    with open(MANIFEST) as fh:
        files = json.load(fh)["files"]
    return all(os.path.exists(os.path.join(STAGED_ROOT, rel)) and _sha256(os.path.join(STAGED_ROOT, rel)) == digest
               for rel, digest in files.items())


if __name__ == "__main__":
    if "--verify" in sys.argv:
        ok = verify()
        print("build_ref: staged copy", "matches its manifest" if ok else "is missing or modified")
        sys.exit(0 if ok else 1)
    sys.exit(0 if stage() else 1)
