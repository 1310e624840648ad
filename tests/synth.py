"""Synthetic inputs now live in the package (bench.py uses them too); kept here as an alias for the tests."""
from soap_bubble_simulation_b200.synth import dt_max, synthetic_field  # noqa: F401
