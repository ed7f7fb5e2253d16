"""`import PG_OP` -> doda_b200.pg_op (B200 replacement of the reference's torch extension of the same name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("doda_b200.pg_op")
