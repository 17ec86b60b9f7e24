"""Pin the oracle to the UNMODIFIED reference and write tests/golden/*.npz --
TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs
/root/reference):

    python -m oracle.make_golden            # from the repo root

Every golden output below is produced by *reference* code (turbopy's
``BorisPush.push``, ``Grid.create_interpolator``, ``Interpolators.interpolate1D``,
``FiniteDifference.*``) imported from /root/reference; the oracle restatement is
asserted bit-identical on the same inputs (and on larger seeded ensembles that
are not stored) before anything is written.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import boris, fd, gather, grid as ogrid, poisson, refload  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

Q_E, M_E = -1.6022e-19, 9.1094e-31          # fixture magnitudes, particle_in_field.py:49-50
Q_P, M_P = 1.6022e-19, 1.6726e-27           # tests/test_computetools.py:26-27
C = 2.9979e8


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def ensemble(rng, n, gamma_kind):
    pos = rng.uniform(0.0, 1.0, (n, 3))
    if gamma_kind == "thermal":              # |p| ~ 0.1 mc  (config 1)
        mom = M_E * C * rng.normal(0.0, 0.1, (n, 3))
    elif gamma_kind == "g10":
        mom = M_E * C * rng.normal(0.0, 10.0, (n, 3))
    else:                                    # beam, gamma log-uniform in [1, 1e3] (config 4)
        g = np.exp(rng.uniform(0.0, np.log(1e3), n))
        bg = np.sqrt(g * g - 1.0)
        mom = np.zeros((n, 3))
        mom[:, 0] = bg * M_E * C
        mom[:, 1:] = 0.01 * mom[:, :1] * rng.normal(0.0, 1.0, (n, 2))
    return pos, mom


def ref_push_steps(tp, pos, mom, q, m, E, B, dt, steps):
    sim = refload.make_owner(tp, clock={"start_time": 0.0, "end_time": dt * steps,
                                        "num_steps": steps})
    # the Clock derives dt by division; force the exact dt under test
    sim.clock.dt = dt
    tool = tp.computetools.BorisPush(sim, {"type": "BorisPush"})
    assert tool.c2 == boris.C2
    for _ in range(steps):
        tool.push(pos, mom, q, m, E, B)


def oracle_soa_steps(pos, mom, q, m, E, B, dt, steps):
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    px, py, pz = (np.ascontiguousarray(mom[:, k]) for k in range(3))
    E = np.asarray(E, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    if E.ndim == 1:
        e = [E[0], E[1], E[2]]
    else:
        e = [np.ascontiguousarray(E[:, k]) for k in range(3)]
    if B.ndim == 1:
        b = [B[0], B[1], B[2]]
    else:
        b = [np.ascontiguousarray(B[:, k]) for k in range(3)]
    for _ in range(steps):
        boris.push_soa(x, y, z, px, py, pz, q, m, *e, *b, dt)
    return np.stack([x, y, z], axis=1), np.stack([px, py, pz], axis=1)


def golden_boris(tp):
    out = {}
    # --- the reference's own known-answer test, tests/test_computetools.py:40-62
    x = np.zeros((1, 3)); p = np.zeros((1, 3))
    E = np.array([[10.0, 0, 0]]); B = np.array([[0, 0, 10.0]])
    ref_push_steps(tp, x, p, Q_P, M_P, E, B, 1e-9, 10)
    assert np.allclose(x, [[2.20028479e-09, -1.04482737e-08, 0]])
    assert np.allclose(p / M_P, [[0.47183691, -1.88168585, 0]])
    out["crossed_x"], out["crossed_p"] = x.copy(), p.copy()

    # --- one-step bitwise vectors, uniform and per-particle fields
    n = 1024
    for kind, seed in (("thermal", 11), ("g10", 12), ("beam", 13)):
        rng = np.random.default_rng(seed)
        pos, mom = ensemble(rng, n, kind)
        Eu = np.array([0.0, 1e5, 0.0]); Bu = np.array([0.0, 0.0, 1.0])
        Ep = rng.normal(0.0, 1e6, (n, 3)); Bp = rng.normal(0.0, 10.0, (n, 3))
        dt = 1e-12
        for tag, E, B in (("uni", Eu, Bu), ("pp", Ep, Bp)):
            rp, rm = pos.copy(), mom.copy()
            ref_push_steps(tp, rp, rm, Q_E, M_E, E, B, dt, 1)
            op, om = oracle_soa_steps(pos, mom, Q_E, M_E, E, B, dt, 1)
            assert bits_equal(rp, op) and bits_equal(rm, om), (kind, tag)
            ap, am = pos.copy(), mom.copy()
            boris.push_aos(ap, am, Q_E, M_E, E, B, dt)
            assert bits_equal(rp, ap) and bits_equal(rm, am), (kind, tag)
            out[f"{kind}_{tag}_pos1"], out[f"{kind}_{tag}_mom1"] = rp, rm
        out[f"{kind}_pos0"], out[f"{kind}_mom0"] = pos, mom
        out[f"{kind}_Ep"], out[f"{kind}_Bp"] = Ep, Bp
    out["Eu"], out["Bu"], out["dt"] = Eu, Bu, np.float64(1e-12)

    # --- config-1 subset: 256 particles, 1000 steps, uniform E x B
    rng = np.random.default_rng(1234)
    pos = rng.uniform(0, 1, (256, 3))
    mom = M_E * C * rng.normal(0, 0.1, (256, 3))
    out["traj_pos0"], out["traj_mom0"] = pos.copy(), mom.copy()
    rp, rm = pos.copy(), mom.copy()
    ref_push_steps(tp, rp, rm, Q_E, M_E, Eu, Bu, 1e-12, 1000)
    op, om = oracle_soa_steps(pos, mom, Q_E, M_E, Eu, Bu, 1e-12, 1000)
    assert bits_equal(rp, op) and bits_equal(rm, om)
    out["traj_pos1000"], out["traj_mom1000"] = rp, rm

    # --- config-4 subset: relativistic beam in strong B, 2000 steps
    rng = np.random.default_rng(4567)
    pos, mom = ensemble(rng, 256, "beam")
    B0 = 10.0
    dt4 = 0.02 * (2 * np.pi * M_E / (abs(Q_E) * B0))
    out["beam_traj_pos0"], out["beam_traj_mom0"] = pos.copy(), mom.copy()
    out["beam_dt"] = np.float64(dt4)
    rp, rm = pos.copy(), mom.copy()
    ref_push_steps(tp, rp, rm, Q_E, M_E, np.zeros(3), np.array([0, 0, B0]), dt4, 2000)
    op, om = oracle_soa_steps(pos, mom, Q_E, M_E, np.zeros(3), np.array([0, 0, B0]), dt4, 2000)
    assert bits_equal(rp, op) and bits_equal(rm, om)
    out["beam_traj_pos2000"], out["beam_traj_mom2000"] = rp, rm

    # --- larger ensembles, checked but not stored (SURVEY 3.2 probe)
    for kind, seed in (("thermal", 21), ("beam", 22)):
        rng = np.random.default_rng(seed)
        n = 200_000
        pos, mom = ensemble(rng, n, kind)
        Ep = rng.normal(0.0, 1e6, (n, 3)); Bp = rng.normal(0.0, 10.0, (n, 3))
        rp, rm = pos.copy(), mom.copy()
        ref_push_steps(tp, rp, rm, Q_E, M_E, Ep, Bp, 1e-12, 20)
        op, om = oracle_soa_steps(pos, mom, Q_E, M_E, Ep, Bp, 1e-12, 20)
        assert bits_equal(rp, op) and bits_equal(rm, om), kind
    np.savez_compressed(os.path.join(GOLDEN, "boris.npz"), **out)
    print("boris.npz:", len(out), "arrays; oracle == reference bit-for-bit")


def golden_gather(tp):
    out = {}
    interp_tool = tp.computetools.Interpolators(tp.Simulation({}), {"type": "Interpolators"})
    for tag, n in (("n30", 30), ("n4097", 4097)):
        sim = refload.make_owner(tp, grid={"N": n, "r_min": 0.0, "r_max": 1.0})
        g = sim.grid
        assert bits_equal(g.r, ogrid.grid_points(0.0, 1.0, n))
        assert g.dr == ogrid.grid_dr(0.0, 1.0, n)
        rng = np.random.default_rng(2345 + n)
        k = 2e8 / 2.998e8
        fields = np.stack([np.cos(2 * np.pi * (c + 1) * k * (g.r - 0.5)) * (1.0 + c)
                           for c in range(6)])                     # (6, n)
        xq = np.concatenate([
            rng.uniform(0.05, 0.95, 1500),
            g.r[[0, 1, n // 2, n - 2, n - 1]],                       # on-node incl. both ends
            np.nextafter(g.r[[1, n // 2, n - 1]], -np.inf),          # 1 ulp below a node
            np.nextafter(g.r[[0, 1, n // 2]], np.inf),               # 1 ulp above a node
            rng.uniform(0.0, g.r[1], 20), rng.uniform(g.r[-2], 1.0, 20),
        ])
        # B: reference call path Interpolators.interpolate1D -> interp1d -> np.interp
        refB = np.stack([interp_tool.interpolate1D(g.r, fields[c])(xq) for c in range(6)])
        # C: same call with 2-D y
        refC = interp_tool.interpolate1D(g.r, fields)(xq)
        for c in range(6):
            ob, st = gather.interp_b(xq, g.r, fields[c])
            assert (st == 0).all() and bits_equal(ob, refB[c]), (tag, "B", c)
        oc, st = gather.interp_c(xq, g.r, fields)
        assert (st == 0).all() and bits_equal(oc, refC), (tag, "C")
        # A: Grid.create_interpolator, one closure per point (reference has no vector form)
        refA = np.full((6, len(xq)), np.nan)
        statA = np.zeros(len(xq), dtype=np.int32)
        for i, x0 in enumerate(xq):
            try:
                f = g.create_interpolator(x0)
                for c in range(6):
                    refA[c, i] = np.squeeze(f(fields[c]))
            except AssertionError:
                statA[i] = gather.BAD_WINDOW
        for c in range(6):
            oa, st = gather.interp_a(xq, g.r, g.dr, fields[c], g.r_min, g.r_max)
            assert np.array_equal(st, statA), (tag, "A status")
            assert bits_equal(oa, refA[c]), (tag, "A", c)
        out[f"{tag}_r"], out[f"{tag}_dr"] = g.r.copy(), np.float64(g.dr)
        out[f"{tag}_fields"], out[f"{tag}_xq"] = fields, xq
        out[f"{tag}_A"], out[f"{tag}_A_status"] = refA, statA
        out[f"{tag}_B"], out[f"{tag}_C"] = refB, refC
    # the reference's own gather test, tests/test_core.py:164-170
    sim = refload.make_owner(tp, grid={"N": 11, "r_min": 0.0, "r_max": 1.0})
    y = sim.grid.r * 10.0
    val = sim.grid.create_interpolator(0.05)(y)
    assert np.allclose(val, 0.5)
    oa, _ = gather.interp_a(np.array([0.05]), sim.grid.r, sim.grid.dr, y)
    assert bits_equal(oa, np.atleast_1d(val))
    # out-of-bounds behaviour: interp1d raises ValueError, create_interpolator asserts
    sim30 = refload.make_owner(tp, grid={"N": 30, "r_min": 0.0, "r_max": 1.0})
    f = interp_tool.interpolate1D(sim30.grid.r, sim30.grid.r)
    for bad in (-1e-9, 1.0 + 1e-9):
        try:
            f(np.array([0.5, bad]))
            raise SystemExit("reference did not raise")
        except ValueError:
            pass
        try:
            sim30.grid.create_interpolator(bad)
            raise SystemExit("reference did not assert")
        except AssertionError:
            pass
    np.savez_compressed(os.path.join(GOLDEN, "gather.npz"), **out)
    print("gather.npz:", len(out), "arrays; oracle == reference bit-for-bit (A, B, C)")


def golden_pif(tp):
    """Gather + EMWave of the particle_in_field example, i.e. what
    tests/fixtures/particle_in_field/output/e_0.5.csv pins: formula A at
    x=0.5 on the N=30 grid for 21 clock times (particle_in_field.py:28-34,45,62)."""
    sim = refload.make_owner(tp, grid={"N": 30, "r_min": 0.0, "r_max": 1.0},
                             clock={"start_time": 0.0, "end_time": 1e-8, "num_steps": 20})
    g, clock = sim.grid, sim.clock
    omega, c_wave = 2e8, 2.998e8
    k = omega / c_wave
    f = g.create_interpolator(0.5)
    times, vals, fields = [], [], []
    for step in range(21):
        t = clock.start_time + clock.dt * step
        E = 1.0 * np.cos(2 * np.pi * (-omega * t + k * (g.r - 0.5)))
        times.append(t); fields.append(E); vals.append(float(np.squeeze(f(E))))
    csv = os.path.join(refload.FIXTURES, "particle_in_field/output/e_0.5.csv")
    ref = np.genfromtxt(csv, delimiter=",")
    # fundamental_cycle diagnoses *before* update (core.py:152-157): CSV row 0 is
    # the initialize() field (t=0) and row i>=1 is the field of time step i-1.
    expect = [vals[0]] + vals[:20]
    assert np.allclose(ref, expect, rtol=1e-5, atol=1e-8)
    # The reference's own point / field diagnostics driven through the same 20-step
    # sequence (diagnose -> update -> advance, core.py:152-158): CSV bytes as goldens
    # for the device-aware diagnostics.
    import tempfile
    csv_bytes = {}
    with tempfile.TemporaryDirectory() as tmp:
        specs = {
            "point": ("point", {"field": "EMField:E", "location": 0.5, "output_type": "csv",
                                "filename": os.path.join(tmp, "point.csv")}),
            "point_node": ("point", {"field": "EMField:E", "location": float(g.r[7]), "output_type": "csv",
                                     "filename": os.path.join(tmp, "point_node.csv")}),
            "field": ("field", {"field": "EMField:E", "component": 0, "output_type": "csv",
                                "filename": os.path.join(tmp, "field.csv")}),
            "field_dump": ("field", {"field": "EMField:E", "component": 0, "output_type": "csv",
                                     "dump_interval": 2e-9, "filename": os.path.join(tmp, "field_dump.csv")}),
            "field_vec": ("field", {"field": "Vec:E", "component": 1, "output_type": "csv",
                                    "filename": os.path.join(tmp, "field_vec.csv")}),
        }
        sim2 = refload.make_owner(tp, grid={"N": 30, "r_min": 0.0, "r_max": 1.0},
                                  clock={"start_time": 0.0, "end_time": 1e-8, "num_steps": 20})
        E = fields[0].copy()
        V = np.stack([fields[0], 2.0 * fields[0], -fields[0]], axis=1).copy()
        diags = {}
        for key, (kind, data) in specs.items():
            d = tp.Diagnostic.lookup(kind)(sim2, data)
            d.inspect_resource({"EMField:E": E, "Vec:E": V})
            d.initialize()
            diags[key] = d
        while sim2.clock.is_running():
            for d in diags.values():
                d.diagnose()
            t = sim2.clock.time
            E[:] = 1.0 * np.cos(2 * np.pi * (-omega * t + k * (g.r - 0.5)))        # EMWave.update
            V[:, 0], V[:, 1], V[:, 2] = E, 2.0 * E, -E
            sim2.clock.advance()
        for key, d in diags.items():
            d.finalize()
            with open(specs[key][1]["filename"], "rb") as fh:
                csv_bytes["csv_" + key] = np.frombuffer(fh.read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "pif.npz"), r=g.r, dr=np.float64(g.dr),
                        times=np.array(times), fields=np.array(fields), e_at_half=np.array(vals), **csv_bytes)
    print("pif.npz: gather A at x=0.5 agrees with the reference's e_0.5.csv golden")


def golden_fd(tp):
    out = {}
    for tag, n, rmax in (("n10", 10, 10.0), ("n1025", 1025, 1.0), ("n1m", 2 ** 20 + 1, 1.0)):
        sim = refload.make_owner(tp, grid={"N": n, "r_min": 0.0, "r_max": rmax,
                                           "coordinate_system": "cylindrical"})
        g = sim.grid
        tool = tp.computetools.FiniteDifference(sim, {"type": "FiniteDifference",
                                                      "method": "centered"})
        assert bits_equal(g.r, ogrid.grid_points(0.0, rmax, n))
        assert bits_equal(g.cell_widths, ogrid.cell_widths(g.r))
        y = np.arange(0, n).astype(np.float64) if n == 10 else \
            np.sin(2 * np.pi * g.r) + 0.1 * g.r ** 2
        rc = tool.centered_difference(y)
        ru = tool.upwind_left(y)
        with np.errstate(divide="ignore"):
            D = tool.del2_radial()
        rd = D @ y
        assert bits_equal(rc, fd.centered_difference(y, g.dr)), tag
        assert bits_equal(ru, fd.upwind_left(y, g.cell_widths)), tag
        assert bits_equal(rd, fd.del2_radial_apply(y, g.r, g.dr)), tag
        if n <= 1025:
            below, diag, above = fd.del2_radial_rows(g.r, g.dr)
            A = D.toarray()
            idx = np.arange(n)
            assert bits_equal(A[idx, idx], diag)
            assert bits_equal(A[idx[1:], idx[:-1]], below[1:])
            assert bits_equal(A[idx[:-1], idx[1:]], above[:-1])
            out[f"{tag}_y"], out[f"{tag}_centered"] = y, rc
            out[f"{tag}_upwind"], out[f"{tag}_del2"] = ru, rd
        else:                                # 1 Mi grid: keep windows only
            sel = np.r_[0:64, n // 2 - 32:n // 2 + 32, n - 64:n]
            out[f"{tag}_sel"] = sel
            out[f"{tag}_centered"], out[f"{tag}_upwind"], out[f"{tag}_del2"] = rc[sel], ru[sel], rd[sel]
        out[f"{tag}_rmax"] = np.float64(rmax)
    np.savez_compressed(os.path.join(GOLDEN, "fd.npz"), **out)
    print("fd.npz:", len(out), "arrays; oracle == reference bit-for-bit")


def golden_fd_operators(tp):
    """Every remaining FiniteDifference matrix builder: stored
    diagonals, dense form and D @ y from the reference; oracle asserted identical."""
    out = {}
    for tag, n, rmax in (("n10", 10, 10.0), ("n1025", 1025, 1.0)):
        sim = refload.make_owner(tp, grid={"N": n, "r_min": 0.0, "r_max": rmax,
                                           "coordinate_system": "cylindrical"})
        g = sim.grid
        tool = tp.computetools.FiniteDifference(sim, {"type": "FiniteDifference", "method": "centered"})
        rng = np.random.default_rng(n)
        y = rng.normal(size=n)
        out[f"{tag}_y"] = y
        for name in fd.BANDED_NAMES:
            with np.errstate(divide="ignore", invalid="ignore"):
                D = getattr(tool, name)()
            data, offsets = fd.banded_operator(name, g.r, g.dr)
            assert list(D.offsets) == list(offsets), name
            assert bits_equal(np.nan_to_num(D.data, nan=-7.0), np.nan_to_num(data, nan=-7.0)), name
            Dy = D @ y
            assert bits_equal(np.nan_to_num(Dy, nan=-7.0), np.nan_to_num(fd.dia_apply(data, offsets, y), nan=-7.0)), name
            # dense form: value equality (scipy builds it through COO and loses the sign of -0.0)
            assert np.array_equal(np.nan_to_num(D.toarray(), nan=-7.0),
                                  np.nan_to_num(fd.dia_toarray(data, offsets, n), nan=-7.0)), name
            out[f"{tag}_{name}_data"], out[f"{tag}_{name}_offsets"] = D.data, np.array(D.offsets)
            out[f"{tag}_{name}_apply"] = Dy
            if n == 10:
                out[f"{tag}_{name}_dense"] = D.toarray()
    np.savez_compressed(os.path.join(GOLDEN, "fd_operators.npz"), **out)
    print("fd_operators.npz:", len(out), "arrays; oracle == reference bit-for-bit")


def golden_poisson(tp):
    """PoissonSolver1DRadial.solve: no reference test exists; the
    oracle is pinned here against the imported reference, bit for bit."""
    out = {}
    for tag, n in (("n64", 64), ("n4097", 4097), ("n1m", (1 << 20) + 1)):
        sim = refload.make_owner(tp, grid={"N": n, "r_min": 0.0, "r_max": 2.0, "coordinate_system": "cylindrical"})
        tool = tp.computetools.PoissonSolver1DRadial(sim, {"type": "PoissonSolver1DRadial"})
        g = sim.grid
        src = np.exp(-((g.r - 0.7) / 0.2) ** 2) - 0.3 * np.cos(5 * g.r)
        with np.errstate(divide="ignore", invalid="ignore"):
            ref = tool.solve(src)
        mine = poisson.solve(src, g.r, g.cell_widths)
        assert bits_equal(ref, mine), tag
        if n <= 4097:
            out[f"{tag}_src"], out[f"{tag}_phi"] = src, ref
        else:
            sel = np.r_[0:32, n // 2 - 16:n // 2 + 16, n - 32:n]
            out[f"{tag}_sel"], out[f"{tag}_phi"] = sel, ref[sel]
            out[f"{tag}_scale"] = np.float64(np.abs(ref).max())
    np.savez_compressed(os.path.join(GOLDEN, "poisson.npz"), **out)
    print("poisson.npz:", len(out), "arrays; oracle == reference bit-for-bit")


def main():
    if not refload.available():
        raise SystemExit("needs /root/reference (build container only)")
    tp = refload.load()
    os.makedirs(GOLDEN, exist_ok=True)
    golden_boris(tp)
    golden_gather(tp)
    golden_pif(tp)
    golden_fd(tp)
    golden_fd_operators(tp)
    golden_poisson(tp)


if __name__ == "__main__":
    main()
