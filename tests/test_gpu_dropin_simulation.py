"""the reference's particle-in-field app through the UNMODIFIED ``turbopy.Simulation.run()``
(turbopy/core.py:127-158) on the GPU box, twice -- stock tools vs ``import turbopy_b200`` -- with byte-identical
CSV output and bit-identical final particle arrays.  The reference comes from oracle/_ref (staged by
oracle/stage_reference.py; hashes pinned by oracle/ref_manifest.json).  Each run is its own process so that the
registry holds exactly one tool set (tests/pif_apps.py)."""
import filecmp
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import refload, stage_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "tests", "pif_apps.py")
CSVS = ["grid.csv", "time.csv", "e_0.5.csv", "particle_p.csv", "particle_x.csv"]

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refload.available(), reason="reference neither at /root/reference nor staged")]


def run_app(tools, app, out, *extra):
    os.makedirs(out, exist_ok=True)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, APP, "--tools", tools, "--app", app, "--out", str(out), *extra],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def same_files(a, b, names):
    for name in names:
        assert filecmp.cmp(os.path.join(a, name), os.path.join(b, name), shallow=False), f"{name} differs"


def same_arrays(a, b):
    for name in ("final_position.npy", "final_momentum.npy"):
        x, y = np.load(os.path.join(a, name)), np.load(os.path.join(b, name))
        assert x.shape == y.shape and np.array_equal(x.view(np.uint64), y.view(np.uint64)), name


def test_staged_reference_is_the_unmodified_reference():
    if refload.kind() == "staged":
        assert stage_reference.verify(), "oracle/_ref does not match the committed manifest of /root/reference"


def test_pif_single_particle_csv_bytes(tmp_path):
    """particle_in_field.toml with ``pusher = "BorisPush"``: one particle, 20 steps, the five CSVs of the app."""
    out = run_app("stock", "single", tmp_path / "stock")
    assert "RUN-OK 20" in out
    out = run_app("b200", "single", tmp_path / "b200")
    assert "RUN-OK 20" in out
    launches = int(out.split("LAUNCHES")[1].split()[0])
    assert launches >= 20, "the B200 run did not launch its kernels"
    same_files(tmp_path / "stock", tmp_path / "b200", CSVS)
    same_arrays(tmp_path / "stock", tmp_path / "b200")
    # and the field diagnostic still matches the reference's own golden (tests/test_pif.py:10-15)
    gold = np.genfromtxt(os.path.join(refload.FIXTURES, "particle_in_field", "output", "e_0.5.csv"), delimiter=",")
    mine = np.genfromtxt(tmp_path / "b200" / "e_0.5.csv", delimiter=",")
    assert np.allclose(gold, mine, rtol=1e-5, atol=1e-8)


def test_pif_as_shipped_raises_the_same_way(tmp_path):
    """The fixture's scalar ``B=0`` makes the stock BorisPush raise in np.cross (computetools.py:436); the drop-in
    must fail with the same exception type instead of silently accepting it."""
    run_app("stock", "single_b0", tmp_path / "stock")
    run_app("b200", "single_b0", tmp_path / "b200")
    a = open(tmp_path / "stock" / "exception.txt").read()
    b = open(tmp_path / "b200" / "exception.txt").read()
    assert a == b == "ValueError\n"


@pytest.mark.parametrize("particles,steps", [(5000, 20), (300_001, 6)])
def test_pif_ensemble_all_paths(tmp_path, particles, steps):
    """N-particle pif: stock Interpolators + BorisPush vs the same two calls on the B200 tools (numpy arrays ->
    host path), vs the fused ``gather_push`` on numpy arrays, vs ``gather_push`` on resident CUDA tensors with the
    device-side particle diagnostic.  All four write the same bytes."""
    extra = ("--particles", str(particles), "--steps", str(steps))
    run_app("stock", "ensemble", tmp_path / "stock", *extra)
    for app in ("ensemble", "ensemble_fused", "ensemble_device"):
        out = run_app("b200", app, tmp_path / app, *extra)
        assert f"RUN-OK {steps}" in out
        assert int(out.split("LAUNCHES")[1].split()[0]) >= steps
        same_files(tmp_path / "stock", tmp_path / app, CSVS)
        same_arrays(tmp_path / "stock", tmp_path / app)
