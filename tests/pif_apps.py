"""The reference's particle-in-field app driven through the UNMODIFIED ``turbopy.Simulation.run()``
(turbopy/core.py:127-158), once with the stock tools and once with ``import turbopy_b200`` -- run as a script in a
subprocess by tests/test_gpu_dropin_simulation.py, which compares the output files byte for byte.

    python tests/pif_apps.py --tools {stock,b200} --app {single,single_b0,ensemble,ensemble_fused,ensemble_device}
                             --out DIR [--particles N] [--steps K]

Everything a Simulation needs comes from the reference: ``Simulation``, ``Grid``, ``SimulationClock``, the stock
``grid`` / ``clock`` / ``point`` diagnostics, and the pif fixture's own ``EMWave`` and ``ParticleDiagnostic``
(tests/fixtures/particle_in_field/particle_in_field.py:12-37,66-99), imported unmodified.  The modules defined here
are the fixture's ``ChargedParticle`` (:40-63) with the one change its stock ``BorisPush`` requires (``B`` as a
3-vector: the fixture's scalar ``B=0`` makes ``np.cross`` raise, app ``single_b0`` checks that both tool sets raise the
same way) and its N-particle generalisation (gather E_y at every particle through the ``Interpolators`` tool, then
``BorisPush.push``), which is what BASELINE configs[1] scales up.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402  (tests may use the oracle package; this is the checker side)


def build_apps(turbopy, pif, tp):
    class ChargedParticleB(pif.ChargedParticle):
        """particle_in_field.py:40-63 with B = zeros(3) (what BorisPush.push needs for np.cross)."""

        def update(self):
            E = np.array([0, float(np.asarray(self.interp_field(self.E)).reshape(-1)[0]), 0])
            self.push(self.position, self.momentum, self.charge, self.mass, E, np.zeros(3))

    class ParticleEnsemble(turbopy.PhysicsModule):
        """N electrons at x uniform in [0.05, 0.95] (seed 2345), p = 0: every step gathers E_y at each particle
        (``Interpolators.interpolate1D(grid.r, E)(x)`` -> numpy.interp rounding) and calls ``BorisPush.push`` with
        per-particle E.  ``mode``: "tools" (the two reference calls), "fused" (turbopy_b200's additive
        ``gather_push``, numpy arrays), "device" (``gather_push`` on resident CUDA tensors)."""

        def __init__(self, owner, input_data):
            super().__init__(owner, input_data)
            n = int(input_data["particles"])
            self.mode = input_data.get("mode", "tools")
            rng = np.random.default_rng(2345)
            pos = np.zeros((n, 3))
            pos[:, 0] = rng.uniform(0.05, 0.95, n)
            mom = np.zeros((n, 3))
            self.charge, self.mass = -1.6022e-19, 9.1094e-31
            self.E = None
            self.pusher = owner.find_tool_by_name("BorisPush")
            self.push = self.pusher.push                                  # bound method captured before initialize()
            self.interp = owner.find_tool_by_name("Interpolators")
            if self.mode == "device":
                import torch
                self.position, self.momentum = tp.to_soa(pos), tp.to_soa(mom)
                self.E_dev = torch.zeros(owner.grid.num_points, dtype=torch.float64, device="cuda")
            else:
                self.position, self.momentum = pos, mom

        def exchange_resources(self):
            self.publish_resource({"ChargedParticle:position": self.position})
            self.publish_resource({"ChargedParticle:momentum": self.momentum})

        def inspect_resource(self, resource):
            if "EMField:E" in resource:
                self.E = resource["EMField:E"]

        def update(self):
            if self.mode == "tools":
                ey = self.interp.interpolate1D(self.owner.grid.r, np.asarray(self.E))(self.position[:, 0])
                E = np.zeros_like(self.position)
                E[:, 1] = ey
                self.push(self.position, self.momentum, self.charge, self.mass, E, np.zeros(3))
            elif self.mode == "fused":
                self.pusher.gather_push(self.position, self.momentum, self.charge, self.mass,
                                        E_grid=(None, np.asarray(self.E), None), B_grid=None)
            else:
                import torch
                self.E_dev.copy_(torch.from_numpy(np.asarray(self.E)))    # the host EMWave module owns the field
                self.pusher.gather_push(self.position, self.momentum, self.charge, self.mass,
                                        E_grid=(None, self.E_dev, None), B_grid=None)

    return ChargedParticleB, ParticleEnsemble


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tools", choices=["stock", "b200"], required=True)
    ap.add_argument("--app", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--particles", type=int, default=5000)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()

    turbopy = refload.load()                                   # the reference, unmodified
    pif = refload.load_pif()
    tp = None
    if a.tools == "b200":
        import turbopy_b200 as tp                              # the import performs the override (core.py:302-312)
        assert turbopy.ComputeTool.lookup("BorisPush") is tp.BorisPush
    else:
        assert turbopy.ComputeTool.lookup("BorisPush").__module__ == "turbopy.computetools"
    ChargedParticleB, ParticleEnsemble = build_apps(turbopy, pif, tp)
    turbopy.PhysicsModule.register("EMWave", pif.EMWave)
    turbopy.PhysicsModule.register("ChargedParticle", pif.ChargedParticle)
    turbopy.PhysicsModule.register("ChargedParticleB", ChargedParticleB)
    turbopy.PhysicsModule.register("ParticleEnsemble", ParticleEnsemble)
    turbopy.Diagnostic.register("ParticleDiagnostic", pif.ParticleDiagnostic)

    particle_diag = "ParticleDiagnostic"
    if a.app == "single":
        modules = {"EMWave": {"amplitude": 1.0, "omega": 2e8}, "ChargedParticleB": {"position": 0.5, "pusher": "BorisPush"}}
    elif a.app == "single_b0":                                 # the fixture as shipped: scalar B = 0
        modules = {"EMWave": {"amplitude": 1.0, "omega": 2e8}, "ChargedParticle": {"position": 0.5, "pusher": "BorisPush"}}
    elif a.app.startswith("ensemble"):
        mode = {"ensemble": "tools", "ensemble_fused": "fused", "ensemble_device": "device"}[a.app]
        modules = {"EMWave": {"amplitude": 1.0, "omega": 2e8}, "ParticleEnsemble": {"particles": a.particles, "mode": mode}}
        if mode == "device":
            particle_diag = "device_particle"                  # the stock CSV utility cannot take CUDA rows
    else:
        raise SystemExit(f"unknown app {a.app}")
    config = {
        "Grid": {"N": 30, "r_min": 0.0, "r_max": 1.0},                                    # particle_in_field.toml:2-10
        "Clock": {"start_time": 0.0, "end_time": 1e-8 * a.steps / 20, "num_steps": a.steps},
        "PhysicsModules": modules,
        "Tools": {"BorisPush": {}, "Interpolators": {}},
        "Diagnostics": {
            "directory": a.out, "output_type": "csv",
            "grid": {"filename": "grid.csv"}, "clock": {"filename": "time.csv"},
            "point": {"field": "EMField:E", "location": 0.5, "filename": "e_0.5.csv"},
            particle_diag: [{"component": "momentum", "filename": "particle_p.csv"},
                            {"component": "position", "filename": "particle_x.csv"}],
        },
    }
    sim = turbopy.Simulation(config)
    try:
        sim.run()                                              # the reference's own loop
    except Exception as exc:                                   # app single_b0: report how it failed
        with open(os.path.join(a.out, "exception.txt"), "w") as f:
            f.write(type(exc).__name__ + "\n")
        print("RAISED", type(exc).__name__, exc)
        return
    tool = sim.find_tool_by_name("BorisPush")
    assert (type(tool).__module__ == "turbopy.computetools") == (a.tools == "stock")
    if hasattr(tool, "check_status"):
        tool.check_status()
    for m in sim.physics_modules:
        if hasattr(m, "position"):
            pos, mom = m.position, m.momentum
            if not isinstance(pos, np.ndarray):
                pos, mom = pos.cpu().numpy(), mom.cpu().numpy()
            np.save(os.path.join(a.out, "final_position.npy"), np.ascontiguousarray(pos))
            np.save(os.path.join(a.out, "final_momentum.npy"), np.ascontiguousarray(mom))
    if tp is not None:
        from turbopy_b200 import _lib
        print("LAUNCHES", _lib.launch_count())
    print("RUN-OK", sim.clock.this_step)


if __name__ == "__main__":
    main()
