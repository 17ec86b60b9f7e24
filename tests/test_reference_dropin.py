"""CPU (needs the reference: /root/reference in the build container, the staged oracle/_ref on the GPU box): the tools drop into the
UNMODIFIED turboPy through its own registry and Simulation plumbing."""
import subprocess
import sys
import os

import pytest

from oracle import refload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference neither at /root/reference nor staged under oracle/_ref")

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r)
from oracle import refload
turbopy = refload.load()                      # the reference, unmodified
stock = turbopy.ComputeTool.lookup("BorisPush")
import turbopy_b200 as tp                     # import performs the override
assert tp.HAVE_TURBOPY
assert tp.ComputeTool is turbopy.ComputeTool and issubclass(tp.BorisPush, turbopy.core.ComputeTool)
for name in ("BorisPush", "Interpolators", "FiniteDifference", "PoissonSolver1DRadial"):
    assert turbopy.ComputeTool.lookup(name) is getattr(tp, name), name
assert turbopy.ComputeTool.lookup("BorisPush") is not stock
assert turbopy.Diagnostic.lookup("point").__module__ == "turbopy.diagnostics"      # stock diagnostics untouched
assert turbopy.Diagnostic.lookup("kinetic_energy") is tp.KineticEnergyDiagnostic
try:
    turbopy.ComputeTool.register("BorisPush", tp.BorisPush)          # no override -> the reference's ValueError
    raise SystemExit("expected ValueError")
except ValueError:
    pass

class Probe(turbopy.PhysicsModule):
    # captures the bound method in __init__, before tool.initialize() (core.py:180-188)
    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self.push = owner.find_tool_by_name("BorisPush").push
        self.ddx = owner.find_tool_by_name("FiniteDifference").setup_ddx()
    def update(self):
        pass
turbopy.PhysicsModule.register("Probe", Probe)
sim = turbopy.Simulation({
    "Grid": {"N": 30, "r_min": 0.0, "r_max": 1.0},
    "Clock": {"start_time": 0.0, "end_time": 1e-8, "num_steps": 20},
    "Tools": {"BorisPush": {}, "Interpolators": {}, "FiniteDifference": {"method": "centered"}},
    "PhysicsModules": {"Probe": {}},
    "Diagnostics": {},
})
sim.run()                                     # the reference's own loop, 20 steps, untouched
tools = {t.name: t for t in sim.compute_tools}
assert type(tools["BorisPush"]) is tp.BorisPush and tools["BorisPush"].owner is sim
assert type(tools["FiniteDifference"]) is tp.FiniteDifference and tools["FiniteDifference"].dr == sim.grid.dr
assert sim.physics_modules[0].push == tools["BorisPush"].push
assert sim.clock.this_step == 20
print("DROPIN-OK")
"""


def test_dropin_into_unmodified_turbopy(tmp_path):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True,
                       cwd=str(tmp_path), env=env)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not refload.tests_dir(), reason="the reference's tests are not present")
def test_reference_suite_substitution_plumbing_and_loud_failure(tmp_path):
    """tests/test_gpu_reference_suite.py runs the reference's own 48 tests against the B200 tools on the GPU box.
    Here, without a device: the substitution takes (names and registry resolve to turbopy_b200), the tests that never
    reach a kernel pass, and every test that does fails with the library's "no CPU fallback" error -- not one of them
    passes on a silent host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the GPU test covers this")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_reference_suite as suite
    rc0, passed0, rep0, tail0 = suite.run_suite(tmp_path, substitute=False)
    assert rc0 == 0 and passed0 == 48, tail0
    rc1, passed1, rep1, tail1 = suite.run_suite(tmp_path, substitute=True)
    assert rc1 != 0 and rep1["launches"] == 0
    assert all(v.startswith("turbopy_b200.") for v in rep1["classes"].values()), rep1
    assert all(v.startswith("turbopy_b200.") for v in rep1["registry"].values()), rep1
    failed = [ln for ln in tail1.splitlines() if ln.startswith("FAILED ")]
    assert passed1 + len(failed) == 48 and len(failed) >= 15, tail1
    assert all("test_computetools.py" in ln and " - Runtime" in ln for ln in failed), failed
    assert "no CPU fallback" in tail1
