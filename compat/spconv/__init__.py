"""`import spconv` -> the B200 engine's spconv-v1.2-compatible surface (put /root/repo/compat on sys.path)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
_impl = importlib.import_module("doda_b200.spconv")
for _sub in ("modules", "conv", "ops", "functional", "tensor"):
    sys.modules["spconv." + _sub] = importlib.import_module("doda_b200.spconv." + _sub)
sys.modules["spconv"] = _impl
