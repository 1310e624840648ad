"""Alias of oracle/oracle_lib.py (the ctypes binding of the CPU oracle) for the tests -- TEST INFRASTRUCTURE ONLY."""
import importlib.util as _u
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "oracle_lib.py")
_spec = _u.spec_from_file_location("_sbx_oracle_lib", _path)
_m = _u.module_from_spec(_spec)
_spec.loader.exec_module(_m)
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
