"""Stage the UNMODIFIED reference where it can travel to the GPU box -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.stage_reference            # build container (has /root/reference)

The reference is pure Python (no compiled artefact to build), and its own ``setup.py`` cannot install it:
``py_modules=['turbopy']`` names a module file that does not exist, so
``pip install --no-index --no-deps --target oracle/_ref /root/reference`` succeeds and installs NOTHING but the
dist-info (tried; recorded in DESIGN.md section 3).  The recipe is therefore a plain byte-for-byte copy of

* ``turbopy/*.py``                                  (the four modules + ``__init__`` / ``__version__``),
* ``tests/fixtures/particle_in_field/``             (the pif app, its TOML input and its golden CSVs),
* ``tests/``                                        (the reference's OWN test-suite, conftest and fixtures, so that
                                                     ``tests/test_gpu_reference_suite.py`` can run it on the GPU box
                                                     with the B200 tools substituted for the stock ones),
* ``LICENSE``                                       (CC0 1.0: copying is permitted),

into ``oracle/_ref/``, which is listed in ``.gitignore`` (build output: it stays out of history) and NOT in
``.gpurunignore`` (so it ships to the GPU box with the snapshot, exactly like the built ``.so``).  No reference source
enters the repository.  ``oracle/ref_manifest.json`` (committed) holds the SHA-256 of every staged file as it is in
``/root/reference``; ``verify()`` re-hashes the staged tree against it, so a GPU-side run can prove that what it
imports is the unmodified reference.

Users: ``oracle/refload.py`` (falls back to ``oracle/_ref`` when ``/root/reference`` is absent), and through it
``tests/`` and ``bench.py``'s CPU arms (``--impl reference`` / ``cpu_baseline``).  The product never imports it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get("TURBOPY_REFERENCE_SOURCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "ref_manifest.json")

# (path below the reference root, path below oracle/_ref)
TREES = [("turbopy", "turbopy"), ("tests/fixtures/particle_in_field", "fixtures/particle_in_field"), ("tests", "tests")]
FILES = [("LICENSE", "LICENSE")]
KEEP = (".py", ".toml", ".csv")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _plan(source):
    plan = []
    for src, dst in TREES:
        base = os.path.join(source, src)
        for root, dirs, files in os.walk(base):
            dirs[:] = sorted(d for d in dirs if d != "__pycache__")
            for name in sorted(files):
                if name.endswith(KEEP):
                    rel = os.path.relpath(os.path.join(root, name), base)
                    plan.append((os.path.join(src, rel), os.path.join(dst, rel)))
    plan += FILES
    return plan


def source_available():
    return os.path.isdir(os.path.join(SOURCE, "turbopy"))


def stage(write_manifest=False, verbose=False):
    """Copy the reference into oracle/_ref.  Returns the manifest {staged path: sha256}."""
    if not source_available():
        raise RuntimeError(f"reference sources not present at {SOURCE}")
    manifest = {}
    for src, dst in _plan(SOURCE):
        s, d = os.path.join(SOURCE, src), os.path.join(DEST, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and _sha(d) == _sha(s)):
            shutil.copyfile(s, d)
        manifest[dst.replace(os.sep, "/")] = _sha(d)
        if verbose:
            print(f"  {src} -> oracle/_ref/{dst}")
    if write_manifest or not os.path.exists(MANIFEST):
        with open(MANIFEST, "w") as f:
            json.dump({"source": "arichar6/turbopy (unmodified)", "files": manifest}, f, indent=1, sort_keys=True)
            f.write("\n")
    return manifest


def staged():
    return os.path.isfile(os.path.join(DEST, "turbopy", "core.py"))


def verify():
    """True when every file of the committed manifest is staged with the recorded hash."""
    if not (staged() and os.path.exists(MANIFEST)):
        return False
    with open(MANIFEST) as f:
        files = json.load(f)["files"]
    for rel, digest in files.items():
        path = os.path.join(DEST, rel)
        if not os.path.exists(path) or _sha(path) != digest:
            return False
    return True


if __name__ == "__main__":
    m = stage(write_manifest="--manifest" in sys.argv, verbose=True)
    print(f"staged {len(m)} files under {DEST}; manifest ok: {verify()}")
