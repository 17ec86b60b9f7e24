"""BASELINE configs[3] at its stated length: the relativistic-beam ensemble pushed 10 000 steps by the UNMODIFIED
reference -- TEST INFRASTRUCTURE ONLY (build container; needs the reference).

    python -m oracle.make_golden_10k            # ~2 min on 8 cores -> tests/golden/beam_10k.npz

Ensemble: 65 536 particles, ``p = gamma beta m c`` along +x with gamma log-uniform in [1, 1e3]
(``default_rng(4567)``) + 1 % transverse spread, ``B = (0, 0, 10 T)``, ``E = 0``,
``dt = 0.02 * 2 pi gamma_min m / (|q| B0)``, 10 000 calls of ``BorisPush.push``
(turbopy/computetools.py:427-441).  Particles are independent, so the ensemble is pushed in shards by forked
workers; every worker runs the reference's own ``push`` on numpy ``(n,3)`` arrays, and the oracle restatement
(``oracle.boris.push_aos``) beside it on the first shard for all 10 000 steps (asserted bit-identical).

Stored (the full final state would be 3 MB): SHA-256 of the final position / momentum bytes of all 65 536 particles
and of the checkpoints at steps 100, 1000 and 5000, plus the final values of every 16th particle.
"""
import hashlib
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import boris, refload  # noqa: E402

Q_E, M_E, C = -1.6022e-19, 9.1094e-31, 2.9979e8
N, STEPS, B0 = 65_536, 10_000, 10.0
CHECKPOINTS = (100, 1000, 5000, 10_000)
DT = 0.02 * (2 * np.pi * M_E / (abs(Q_E) * B0))


def beam_ensemble(n=N, seed=4567):
    """Identical to tests/helpers.beam_ensemble (kept in step by tests/test_oracle_goldens.py)."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.0, 1.0, (n, 3))
    g = np.exp(rng.uniform(0.0, np.log(1e3), n))
    bg = np.sqrt(g * g - 1.0)
    mom = np.zeros((n, 3))
    mom[:, 0] = bg * M_E * C
    mom[:, 1:] = 0.01 * mom[:, :1] * rng.normal(0.0, 1.0, (n, 2))
    return pos, mom


def _worker(args):
    lo, hi, with_oracle = args
    tp = refload.load()
    sim = refload.make_owner(tp, clock={"start_time": 0.0, "end_time": DT * STEPS, "num_steps": STEPS})
    sim.clock.dt = DT                       # the Clock derives dt by division; force the exact dt under test
    tool = tp.computetools.BorisPush(sim, {"type": "BorisPush"})
    pos, mom = beam_ensemble()
    pos, mom = pos[lo:hi].copy(), mom[lo:hi].copy()
    op, om = pos.copy(), mom.copy()
    E, B = np.zeros(3), np.array([0.0, 0.0, B0])
    snaps = {}
    for s in range(1, STEPS + 1):
        tool.push(pos, mom, Q_E, M_E, E, B)
        if with_oracle:
            boris.push_aos(op, om, Q_E, M_E, E, B, DT)
        if s in CHECKPOINTS:
            snaps[s] = (pos.copy(), mom.copy())
            if with_oracle:
                assert np.array_equal(op.view(np.uint64), pos.view(np.uint64)), f"oracle != reference at step {s}"
                assert np.array_equal(om.view(np.uint64), mom.view(np.uint64)), f"oracle != reference at step {s}"
    return lo, snaps


def main():
    if not refload.available():
        raise SystemExit("reference not present")
    workers = min(os.cpu_count() or 1, 16)
    edges = np.linspace(0, N, workers + 1).astype(int)
    jobs = [(int(edges[i]), int(edges[i + 1]), i == 0) for i in range(workers)]
    with mp.get_context("fork").Pool(workers) as pool:
        parts = sorted(pool.map(_worker, jobs))
    out = {"n": np.int64(N), "steps": np.int64(STEPS), "dt": np.float64(DT), "B0": np.float64(B0),
           "checkpoints": np.array(CHECKPOINTS, dtype=np.int64), "stride": np.int64(16)}
    for s in CHECKPOINTS:
        pos = np.concatenate([snaps[s][0] for _, snaps in parts])
        mom = np.concatenate([snaps[s][1] for _, snaps in parts])
        out[f"sha_pos_{s}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pos).tobytes()).digest(), dtype=np.uint8)
        out[f"sha_mom_{s}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(mom).tobytes()).digest(), dtype=np.uint8)
    out["pos_final_16"], out["mom_final_16"] = pos[::16].copy(), mom[::16].copy()
    p0, m0 = beam_ensemble()
    # |p| is conserved by a pure rotation up to rounding: a sanity check on the whole run
    drift = np.abs(np.sqrt((mom ** 2).sum(1)) / np.sqrt((m0 ** 2).sum(1)) - 1).max()
    assert drift < 1e-10, drift
    path = os.path.join(ROOT, "tests", "golden", "beam_10k.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {N} particles x {STEPS} steps by the reference; oracle == reference on shard 0; "
          f"max |p| drift {drift:.2e}")


if __name__ == "__main__":
    main()
