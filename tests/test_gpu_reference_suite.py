"""The reference's OWN test-suite (tests/test_computetools.py, test_core.py, test_diagnostics.py, test_pif.py,
test_bos.py -- 48 tests, unmodified, from /root/reference or the staged oracle/_ref/tests) run on the GPU box with
the B200 tools substituted for the stock ``BorisPush`` / ``Interpolators`` / ``FiniteDifference`` /
``PoissonSolver1DRadial`` (tests/b200_substitute.py rebinds the names the test modules star-import and takes the
registry).  Every assertion the reference makes about those tools -- the crossed-field known answer
(test_computetools.py:40-62), free streaming (:15-37), ``interpolate1D`` incl. the 'quadratic' kind (:71-86), the
stencils (:104-138), exact element equality of every operator builder's ``toarray()`` (:141-291) -- then holds for the
CUDA path, through the tools' host-array entry points (numpy in, numpy out, in-place push), which is how a turboPy
user calls them.  The same suite is run with the stock tools first: the pass counts must agree."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

from oracle import refload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (refload.available() and refload.tests_dir()),
                                 reason="the reference's tests are neither at /root/reference nor staged")]


def run_suite(tmp_path, substitute):
    work = tmp_path / ("b200" if substitute else "stock")
    shutil.copytree(refload.tests_dir(), work / "tests",              # a copy: the suite writes ./tmp/ and ./output/
                    ignore=shutil.ignore_patterns("__pycache__", "tmp"))
    report = work / "report.json"
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", B200_SUBSTITUTE="1" if substitute else "0",
               B200_SUBSTITUTE_REPORT=str(report),
               PYTHONPATH=os.pathsep.join([*refload.paths(), ROOT, os.path.join(ROOT, "tests")]))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-q", "-p", "b200_substitute", "-p", "no:cacheprovider",
                        "-W", "ignore", "--rootdir", str(work), "-c", os.devnull],
                       capture_output=True, text=True, env=env, cwd=work)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert report.exists(), tail
    with open(report) as f:
        rep = json.load(f)
    m = re.search(r"(\d+) passed", r.stdout)
    return r.returncode, int(m.group(1)) if m else 0, rep, tail


def test_reference_suite_passes_with_b200_tools(tmp_path):
    rc0, passed0, rep0, tail0 = run_suite(tmp_path, substitute=False)
    assert rc0 == 0 and passed0 >= 48, tail0                           # 48 passed with the stock tools
    assert all(v.startswith("turbopy.computetools.") for v in rep0["classes"].values()), rep0

    rc1, passed1, rep1, tail1 = run_suite(tmp_path, substitute=True)
    assert rc1 == 0, tail1
    assert passed1 == passed0, (passed0, passed1, tail1)
    assert all(v.startswith("turbopy_b200.") for v in rep1["classes"].values()), rep1
    assert all(v.startswith("turbopy_b200.") for v in rep1["registry"].values()), rep1
    assert rep1["launches"] and rep1["launches"] >= 20, f"the substituted run launched no kernels: {rep1}"
