"""TEST INFRASTRUCTURE ONLY -- stand-in for the `qtoml` package the reference
imports at turbopy/constructors.py:2 (not installed here, no network).
Backed by the standard library's tomllib."""
import tomllib


def load(f):
    s = f.read()
    return tomllib.loads(s if isinstance(s, str) else s.decode())


loads = tomllib.loads
