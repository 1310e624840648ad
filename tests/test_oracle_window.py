"""The oracle's window mode (a latitude slab of a larger sphere, used to check sharded runs at sizes the whole-sphere
oracle cannot finish in seconds) against the whole-sphere oracle: bit-identical away from the cut edges."""
import numpy as np
import pytest

import oracle_lib as O
from synth import dt_max, synthetic_field


@pytest.mark.parametrize("nt,nphi,steps,dt", [(200, 64, 30, None), (128, 256, 140, 5e-3), (61, 50, 25, None)])
def test_window_equals_whole_sphere_inside(nt, nphi, steps, dt):
    f = synthetic_field(nt, nphi)
    if dt is None:
        dt = 0.25 * dt_max(nt, nphi)
    full = O.OracleDrainage(nt, nphi)
    full.data[:] = f
    full.step(dt, steps)
    if dt == 5e-3:
        assert full.active_tfe_count() >= 1  # fronts alive: the window applies them to its own rows only
    third = nt // 3
    for r0, rows in [(0, third + steps + 3), (third - steps - 2, third + 2 * steps + 4), (nt - third - steps - 3, third + steps + 3)]:
        r0 = max(r0, 0)
        rows = min(rows, nt - r0)
        w = O.OracleDrainage(nt, nphi, window=(r0, rows))
        w.data[:] = f[r0:r0 + rows]
        w.step(dt, steps)
        lo = 0 if r0 == 0 else steps
        hi = rows if r0 + rows == nt else rows - steps
        assert hi > lo
        assert np.array_equal(w.data[lo:hi], full.data[r0 + lo:r0 + hi])
        assert w.fronts() == full.fronts() and w.current_time() == full.current_time()


def test_window_rejects_bad_ranges():
    with pytest.raises(ValueError):
        O.OracleDrainage(64, 32, window=(60, 10))
    with pytest.raises(ValueError):
        O.OracleDrainage(64, 32, window=(3, 1))
