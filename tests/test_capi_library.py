"""CPU: the C-ABI library builds, loads, exports every symbol the header
declares, and refuses to compute without a B200 (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from turbopy_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib


def header_functions():
    text = open(os.path.join(ROOT, "include", "turbopy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tp_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = header_functions()
    assert len(declared) >= 20
    assert sorted(lib.SIGNATURES) == declared


def test_every_declared_symbol_is_exported(lib):
    handle = C.CDLL(lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(handle, name), f"{name} declared in include/turbopy_b200.h but not exported"


def test_version_and_error_strings(lib):
    l = lib.lib()
    assert l.tp_version() == 100
    assert lib.error_string(0) == "ok"
    assert "no CPU fallback" in lib.error_string(lib.TP_E_NODEVICE)
    assert lib.error_string(-2).startswith("turbopy_b200")


def test_struct_layouts_match_header(lib):
    assert C.sizeof(lib.Particles) == 7 * 8
    assert C.sizeof(lib.PushConsts) == 4 * 8
    assert C.sizeof(lib.GridFields) == 5 * 8 + 6 * 8 + 6 * 8
    assert C.sizeof(lib.Linspace) == 4 * 8


def test_sass_is_sm100_with_tma(lib):
    """The shipped cubin is sm_100a and the staging / particle pipeline really
    use the TMA (UBLKCP = cp.async.bulk) and mbarrier (SYNCS) instructions."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out and "SYNCS" in out
    assert "DFMA" in out                      # fp64 pipe, not tensor cores


def test_streaming_kernels_fit_the_ctas_they_launch_per_sm(lib):
    """The register-path streaming kernels launch a fixed number of 256-thread CTAs per SM (csrc/fd_kernels.cu table;
    the Poisson passes take theirs from the occupancy query but were tuned at four): that many must be RESIDENT, or
    the surplus runs as a ragged second wave.  fd_centered_kernel once grew from 48 to 52 registers through an
    unrelated header change and dropped from 98 % to 72 % of the copy peak unnoticed -- this pins the budgets."""
    import re
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-res-usage", lib.LIB_PATH], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+?):\s*\n\s*REG:(\d+)", out))
    assert regs, "cuobjdump -res-usage printed no functions"
    resident = {"fd_centered_kernel": 5, "fd_upwind_kernel": 4, "fd_del2_radial_kernel": 4, "fd_dia_kernel": 5,
                "fd_dia_const_kernel": 4, "fd_dia_const_t_kernel": 4, "moments_partial_kernel": 4,
                "poisson_sums_kernel": 4, "poisson_apply_kernel": 4}
    seen = set()
    for name, r in regs.items():
        for key, ctas in resident.items():
            if re.search(rf"\d+{key}(I|E)", name):                    # mangled: <len><name> then template args or params
                seen.add(key)
                alloc = -(-int(r) // 8) * 8                            # registers are allocated in units of 8 per thread
                assert alloc * 256 * ctas <= 65536, f"{name}: {r} registers, {ctas} CTAs of 256 threads do not fit one SM"
    assert seen == set(resident), f"kernels not found in the library: {set(resident) - seen}"


def test_fused_diagnostic_instances_are_in_the_library(lib):
    """The MOM instances (kernel 3 fused into kernel 1) the launchers select for tp_gather_push_moments_f64 /
    tp_boris_push_moments_f64: staged grid (generic and E_y-only), per-tile windows, uniform-field push -- formula B
    (= 2), SoA, and they fit the one-CTA-per-SM ring (512 threads: at most 128 registers, no large spill frame)."""
    import re
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-res-usage", lib.LIB_PATH], capture_output=True, text=True).stdout
    usage = {m.group(1): (int(m.group(2)), int(m.group(3)))
             for m in re.finditer(r"Function (\S+?):\s*\n\s*REG:(\d+) STACK:(\d+)", out)}
    wanted = ["22gather_push_tma_kernelILi2ELi1ELb0ELj0ELb1EE", "22gather_push_tma_kernelILi2ELi1ELb0ELj2ELb1EE",
              "22gather_push_win_kernelILi2ELb0ELb1EE", "15push_tma_kernelILb0ELb1EE"]
    for key in wanted:
        hits = [(name, v) for name, v in usage.items() if key in name]
        assert len(hits) == 1, f"{key}: {len(hits)} instances in the library"
        regs, stack = hits[0][1]
        assert regs <= 128 and stack <= 256, f"{hits[0][0]}: {regs} registers, {stack} B of stack"


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import turbopy_b200 as tp
    from helpers import make_owner
    l = lib.lib()
    assert l.tp_device_info(None, None, None, None, None) == lib.TP_E_NODEVICE
    tool = tp.BorisPush(make_owner(), {"type": "BorisPush"})
    pos = np.zeros((4, 3)); mom = np.zeros((4, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tool.push(pos, mom, 1.0, 1.0, np.zeros(3), np.zeros(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tool.gather_push(pos, mom, 1.0, 1.0, E_grid=None, B_grid=None)
    fdt = tp.FiniteDifference(make_owner(), {"type": "FiniteDifference", "method": "centered"})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fdt.centered_difference(np.zeros(10))
    assert np.array_equal(pos, np.zeros((4, 3)))
