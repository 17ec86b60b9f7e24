"""TEST INFRASTRUCTURE ONLY -- restores the numpy aliases the reference still
uses (turbopy/core.py:531,633; tests/test_computetools.py:29-30)."""
import numpy as np

if not hasattr(np, "int"):
    np.int = int
if not hasattr(np, "float"):
    np.float = float
