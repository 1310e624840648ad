"""Reference vectors (scripts/make_reference_vectors.rs -> tests/golden/drainage_*.bin, interference_*.bin).

When vector files written by the real reference are present, the C oracle must reproduce every checkpoint BIT FOR BIT
(same operation order, no FMA contraction, the same libm) -- that is what pins the drainage half of the oracle.  They
cannot be produced in this image (no cargo / rustc); until somebody runs the generator these tests skip and the oracle's
header keeps saying "parity unpinned" for drainage.  The file format itself is always tested (written from the oracle
into a temporary directory, never under tests/golden/)."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O
from reference_vectors import (DrainageVector, read_interference, read_vector, write_interference, write_vector)

HERE = os.path.dirname(os.path.abspath(__file__))
DRAINAGE = sorted(glob.glob(os.path.join(HERE, "golden", "drainage_*.bin")))
INTERF = sorted(glob.glob(os.path.join(HERE, "golden", "interference_*.bin")))
NEED = "no reference vectors under tests/golden/ (run scripts/make_reference_vectors.rs with cargo)"


def oracle_run(v: DrainageVector):
    ora = O.OracleDrainage(v.num_theta, v.num_phi, **v.oracle_kwargs())
    ora.set_marginal_regeneration(v.tfe_enabled)
    done = 0
    for steps, t, cnt, fld in v.checkpoints:
        ora.step(v.dt, steps - done)
        done = steps
        yield steps, t, cnt, fld, ora


@pytest.mark.skipif(not DRAINAGE, reason=NEED)
@pytest.mark.parametrize("path", DRAINAGE or ["-"])
def test_oracle_reproduces_reference_drainage_vectors(path):
    v = read_vector(path)
    for steps, t, cnt, fld, ora in oracle_run(v):
        assert ora.current_time() == t, (path, steps)
        assert ora.active_tfe_count() == cnt, (path, steps)
        assert np.array_equal(ora.data, fld), (path, steps, float(np.abs(ora.data - fld).max()))


@pytest.mark.skipif(not INTERF, reason=NEED)
@pytest.mark.parametrize("path", INTERF or ["-"])
def test_oracle_reproduces_reference_interference_vectors(path):
    n_film, pts = read_interference(path)
    calc = O.OracleInterference(n_film)
    for d, c, r, g, b, lr, lg, lb in pts:
        assert tuple(calc.color(d, c)) == (r, g, b), (d, c)
        assert tuple(calc.color_linear(d, c)) == (lr, lg, lb), (d, c)


def test_vector_format_round_trip(tmp_path):
    """What the Rust generator writes is what the reader reads: same layout produced here from the oracle."""
    v = DrainageVector(16, 16, 32, 500.0, 0.05, 10.0, 1000.0, 0.001, 1e-9, 9.81, 5e-3, True)
    ora = O.OracleDrainage(16, 32, **v.oracle_kwargs())
    done = 0
    for steps in (1, 150, 400):
        ora.step(v.dt, steps - done)
        done = steps
        v.checkpoints.append((steps, ora.current_time(), ora.active_tfe_count(), np.array(ora.data)))
    p = str(tmp_path / "oracle_written_drainage.bin")
    write_vector(p, v)
    assert os.path.getsize(p) == 92 + 3 * (24 + 16 * 32 * 8)  # the header and record sizes the Rust side writes
    w = read_vector(p)
    assert (w.num_theta, w.num_phi, w.dt, w.tfe_enabled, w.diameter) == (16, 32, 5e-3, True, 0.05)
    assert np.array_equal(w.initial, np.full((16, 32), 500.0 * 1e-9))
    for (s0, t0, c0, f0), (s1, t1, c1, f1, o) in zip(v.checkpoints, oracle_run(w)):
        assert (s0, t0, c0) == (s1, t1, c1) and np.array_equal(f0, f1) and np.array_equal(o.data, f0)
    pts = np.array([[500.0, 1.0, 0.3, 0.0, 0.3, 0.07, 0.0, 0.07]])
    q = str(tmp_path / "oracle_written_interference.bin")
    write_interference(q, 1.33, pts)
    n, back = read_interference(q)
    assert n == 1.33 and np.array_equal(back, pts)
    with pytest.raises(ValueError):
        read_vector(q)
