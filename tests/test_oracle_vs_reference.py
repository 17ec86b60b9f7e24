"""CPU (needs the reference: /root/reference in the build container, the staged oracle/_ref on the GPU box): the oracle restatement against the UNMODIFIED reference on
seeded random inputs including the awkward ones (non-finite values, on-node and end points, NaN abscissas,
repeated values).  Runs in a subprocess so that the stock tools are the ones registered (importing turbopy_b200
overrides them).  make_golden.py pins the oracle once when the goldens are generated; this keeps it pinned."""
import os
import subprocess
import sys

import pytest

from oracle import refload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference neither at /root/reference nor staged under oracle/_ref")

SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, %(root)r)
from oracle import refload, boris as oboris, gather as ogather, fd as ofd, poisson as opoisson
turbopy = refload.load()                                  # the reference, unmodified; stock tools registered


def same(a, b, what):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    nan = np.isnan(a)
    assert np.array_equal(nan, np.isnan(b)), what + ": NaN pattern"
    assert np.array_equal(a[~nan].view(np.uint64), b[~nan].view(np.uint64)), what + ": bits"


rng = np.random.default_rng(2024)
Q, M = -1.6022e-19, 9.1094e-31
specials = [np.inf, -np.inf, np.nan, 1e308, -1e308, 5e-324, 0.0, -0.0, 1e-200]

# ---- BorisPush.push: random ensembles over 9 decades of momentum, uniform and per-particle fields, non-finite entries
for trial in range(12):
    n = int(rng.integers(1, 400))
    dt = 10.0 ** rng.uniform(-14, -9)
    sim = refload.make_owner(turbopy, clock={"start_time": 0.0, "end_time": dt * 10, "num_steps": 10})
    dt = sim.clock.dt                                      # (end - start) / num_steps, not necessarily the dt above
    tool = turbopy.ComputeTool.lookup("BorisPush")(sim, {"type": "BorisPush"})
    pos = rng.uniform(-1, 1, (n, 3)); mom = rng.normal(0, 10.0 ** rng.uniform(-3, 3) * M * 2.9979e8, (n, 3))
    pp = trial %% 2 == 1
    E = rng.normal(0, 1e6, (n, 3) if pp else 3); B = rng.normal(0, 5.0, (n, 3) if pp else 3)
    if trial >= 6:
        for k in range(min(n, 24)):
            tgt = (mom, pos, E, B)[k %% 4] if pp else (mom, pos)[k %% 2]
            tgt[rng.integers(0, n), rng.integers(0, 3)] = specials[k %% len(specials)]
    rp, rm = pos.copy(), mom.copy()
    with np.errstate(all="ignore"):
        for _ in range(3):
            tool.push(pos, mom, Q, M, E, B)
            oboris.push_aos(rp, rm, Q, M, E, B, dt)
    same(rp, pos, f"push trial {trial}: position"); same(rm, mom, f"push trial {trial}: momentum")

# ---- Interpolators.interpolate1D (numpy.interp for 1-D y, scipy's N-D path for 2-D y) and Grid.create_interpolator
interp = turbopy.ComputeTool.lookup("Interpolators")(refload.make_owner(turbopy), {"type": "Interpolators"})
for trial in range(12):
    ng = int(rng.choice([2, 3, 11, 30, 257, 1025]))
    lo, hi = float(rng.uniform(-3, 0)), float(rng.uniform(0.1, 4))
    sim = refload.make_owner(turbopy, grid={"N": ng, "r_min": lo, "r_max": hi})
    r = sim.grid.r
    xp = r if trial %% 3 else np.sort(np.concatenate([[lo, hi], rng.uniform(lo, hi, ng - 2)]))
    if len(np.unique(xp)) != len(xp):
        xp = r
    lo, hi = float(xp[0]), float(xp[-1])                   # the grid's own end points (r_max is rounded into r[-1])
    y = np.cos(3 * xp) * 10.0 ** rng.uniform(-2, 5)
    if trial %% 4 == 3 and ng > 3:
        y[rng.integers(0, ng)] = np.inf                    # numpy.interp's NaN fallbacks
        y[rng.integers(0, ng)] = -np.inf
    xq = np.concatenate([rng.uniform(lo, hi, 300), xp[rng.integers(0, ng, 40)], [lo, hi, np.nan],
                         np.nextafter(xp[rng.integers(0, ng, 20)], hi).clip(lo, hi),
                         np.nextafter(xp[rng.integers(0, ng, 20)], lo).clip(lo, hi)])
    with np.errstate(all="ignore"):
        out, st = ogather.interp_b(xq, xp, y)
        assert (st == 0).all()
        same(out, interp.interpolate1D(xp, y)(xq), f"interp B trial {trial}")
        y2 = np.stack([y, -2 * y, y * y])
        out, st = ogather.interp_c(xq, xp, y2)
        assert (st == 0).all()
        same(out, interp.interpolate1D(xp, y2)(xq), f"interp C trial {trial}")
    for bad, code in ((lo - 1e-9 * (hi - lo), ogather.BELOW), (hi + 1e-9 * (hi - lo), ogather.ABOVE)):
        _, st = ogather.interp_b(np.array([bad]), xp, y)
        assert st[0] == code
        try:
            interp.interpolate1D(xp, y)(np.array([bad]))
            raise SystemExit("interp1d accepted an out-of-range point")
        except ValueError:
            pass
    # formula A: per point through the reference's closure; asserts <-> status
    yf = np.cos(3 * r)
    pts = np.concatenate([rng.uniform(lo, hi, 60), r[rng.integers(0, ng, 12)], [lo, hi, lo - 1.0, hi + 1.0, np.nan]])
    vals, st = ogather.interp_a(pts, r, sim.grid.dr, yf, sim.grid.r_min, sim.grid.r_max)
    for x, v, s in zip(pts, vals, st):
        try:
            ref = float(np.squeeze(sim.grid.create_interpolator(x)(yf)))
        except AssertionError:
            assert s != 0, (x, s)
            continue
        assert s == 0, (x, s)
        same(np.array([v]), np.array([ref]), f"interp A at {x!r}")

# ---- FiniteDifference: stencils and every matrix builder, incl. non-finite operands
for trial in range(8):
    n = int(rng.choice([4, 5, 10, 33, 1025]))
    sim = refload.make_owner(turbopy, grid={"N": n, "r_min": 0.0, "r_max": float(10.0 ** rng.uniform(-1, 2)),
                                            "coordinate_system": "cylindrical"})
    g = sim.grid
    fdt = turbopy.ComputeTool.lookup("FiniteDifference")(sim, {"type": "FiniteDifference", "method": "centered"})
    y = rng.normal(size=n) * 10.0 ** rng.uniform(-3, 3)
    if trial >= 4:
        y[rng.integers(0, n)] = specials[trial %% 3]
    with np.errstate(all="ignore"):
        same(ofd.centered_difference(y, g.dr), fdt.centered_difference(y), "centered")
        same(ofd.upwind_left(y, g.cell_widths), fdt.upwind_left(y), "upwind")
        same(ofd.del2_radial_apply(y, g.r, g.dr), fdt.del2_radial() @ y, "del2_radial")
        for name in ofd.BANDED_NAMES:
            data, offsets = ofd.banded_operator(name, g.r, g.dr)
            D = getattr(fdt, name)()
            assert list(D.offsets) == list(offsets), name
            same(data, D.data, name + ": stored diagonals")
            same(ofd.dia_apply(data, offsets, y), D @ y, name + " @ y")

# ---- PoissonSolver1DRadial.solve
for n in (3, 4, 64, 1000):
    sim = refload.make_owner(turbopy, grid={"N": n, "r_min": 0.0, "r_max": 2.0, "coordinate_system": "cylindrical"})
    ps = turbopy.ComputeTool.lookup("PoissonSolver1DRadial")(sim, {"type": "PoissonSolver1DRadial"})
    src = rng.normal(size=n)
    with np.errstate(all="ignore"):
        same(opoisson.solve(src, sim.grid.r, sim.grid.cell_widths), ps.solve(src), f"poisson n={n}")
# ---- the reference's Grid.r IS the linspace formula the *_lin_* kernels rebuild in registers
import turbopy_b200 as tp                                  # (auto-registration disabled by the environment)
for trial in range(40):
    n = int(rng.integers(2, 3000))
    lo = float(rng.uniform(-5, 5)); hi = lo + float(10.0 ** rng.uniform(-4, 3))
    sim = refload.make_owner(turbopy, grid={"N": n, "r_min": lo, "r_max": hi})
    lin = tp.GridOnDevice.probe_linspace(np.ascontiguousarray(sim.grid.r), (sim.grid.r_min, sim.grid.r_max))
    assert lin is not None and lin.n == n, (n, lo, hi)
    i = np.arange(n, dtype=np.float64); t = i * lin.step; t[-1] = 1.0
    same(lin.start + lin.span * t, sim.grid.r, "linspace formula vs Grid.r")
    same(sim.grid.r[1:] - sim.grid.r[:-1], sim.grid.cell_widths, "cell_widths = r[1:] - r[:-1]")
assert tp.GridOnDevice.probe_linspace(np.linspace(0.0, 1.0, 50) ** 2) is None
print("ORACLE-PINNED")
"""


def test_oracle_matches_the_unmodified_reference_on_random_and_awkward_inputs():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", TURBOPY_B200_NO_AUTOREGISTER="1")
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, env=env,
                         timeout=600)
    assert out.returncode == 0 and "ORACLE-PINNED" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
