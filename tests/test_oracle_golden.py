"""The oracle is pinned: its output equals the unmodified reference's (tests/golden/, made by make_golden.py from
oracle/_ref) and the known answers the reference quotes itself."""
import os
import tempfile

import numpy as np
import pytest

import oracle_lib as O
from fixtures import golden, golden_lines, ring_scenario, sha256_lines, write_ring_files


def test_ring_readme_lines_in_golden():
    g = golden("ring")
    assert g["count"] == 38 and len(g["lines"]) == 38
    assert g["sha256"] == "9f3761e80125fd0989022ab985331e0c07c58884f8ce7fbd6e3becf58d196550"
    for l in g["readme_lines"]:                      # reference traffic/README.md:121-126
        assert l in g["lines"]
    for l in ("RE,0,9,370,0,411,END", "RE,1,6,625,7,666,END", "RE,2,8,340,9,381,END"):
        assert l in g["lines"]


@pytest.mark.parametrize("mode", ["timewarp1", "timewarp4", "closure", "sequential"])
def test_oracle_ring_files(mode):
    g = golden("ring")
    with tempfile.TemporaryDirectory() as td:
        rd, trip, _ = write_ring_files(td)
        s = O.Sim.traffic_files(rd, trip)
        _run(s, mode, 5)
        lines = O.records_to_lines(s.committed())
    assert sorted(lines) == g["lines"]
    assert s.digest() == g["digest"]


def test_oracle_ring_arrays_closure():
    sc, g = ring_scenario()
    s = O.Sim.traffic_scenario(sc)
    s.run_closure(2)
    assert sorted(O.records_to_lines(s.committed())) == g["lines"]


def _run(s, mode, switch):
    if mode == "timewarp1":
        s.run_timewarp(1, 100, switch)
    elif mode == "timewarp4":
        s.run_timewarp(4, 100, switch)
    elif mode == "closure":
        s.run_closure(3)
    else:
        s.run_sequential()


@pytest.mark.parametrize("mode", ["timewarp1", "timewarp4", "closure", "sequential"])
def test_oracle_phold_100(mode):
    g = golden("phold_100_1600")
    assert g["count"] == 7517
    assert g["sha256"] == "2d7ee1e556d199c3b051491d259cab8c6e161ec2197811d03f2cdea9b3fff3a1"
    s = O.Sim.phold(100, 1600, 0.1, 1.0)
    n = _run(s, mode, 1)
    lines = O.records_to_lines(s.committed())
    assert len(lines) == 7517
    assert sorted(lines) == golden_lines("phold_100_1600")
    assert s.digest() == g["digest"]
    if mode == "timewarp4":
        assert s.stats()["cancels"] > 0      # the emulation really rolls back
    for l in ("RE,900,0,519,0,519,END", "RE,1500,0,1107,0,1107,END", "RE,300,0,4230,0,4230,END"):
        assert l in lines


@pytest.mark.parametrize("name", ["phold_64_1024", "phold_1000_4000"])
@pytest.mark.parametrize("mode", ["timewarp4", "closure"])
def test_oracle_phold_shapes(name, mode):
    g = golden(name)
    n_lp, n_init, ratio, lam = g["args"]
    s = O.Sim.phold(int(n_lp), int(n_init), ratio, lam)
    _run(s, mode, 1)
    assert s.digest() == g["digest"]
    assert sha256_lines(O.records_to_lines(s.committed())) == g["sha256"]


def test_oracle_phold_10000_digest():
    g = golden("phold_10000_160000")
    assert g["count"] == 728183
    s = O.Sim.phold(10000, 160000, 0.1, 1.0)
    s.run_closure(4, keep_records=False)
    assert s.digest() == g["digest"]
    s.run_sequential(keep_records=False)
    assert s.digest() == g["digest"]


def test_oracle_phold_tables_pins():
    """libstdc++ stream pins of the PHOLD tables (seed 1, lambda 1.0, ratio 0.1)."""
    n = 10_000_000
    lat = np.empty(n, dtype=np.int64)
    rem = np.empty(n, dtype=np.int32)
    O.lib().two_phold_tables(1, 1.0, 0.1, n, O._p(lat), O._p(rem))
    assert lat[:10].tolist() == [14103, 61368, 24712, 113589, 272865, 73275, 3518, 75438, 772, 6918]
    assert rem[:10].tolist() == [0, 0, 0, 0, 0, 0, 1, 0, 1, 1]
    assert int(lat.max()) == 1570160
    assert abs(float(lat.mean()) - 99939.3) < 0.1
    O.lib().two_fnv_fold.restype = __import__("ctypes").c_uint64
    O.lib().two_fnv_fold.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_int64]
    assert O.lib().two_fnv_fold(O._p(lat), n) == 0x0933b478f8e84063


def test_stateful_sequential_differs_from_stateless():
    a = O.Sim.phold(100, 1600, 0.1, 1.0)
    a.run_sequential()
    b = O.Sim.phold(100, 1600, 0.1, 1.0, stateful=True)
    b.run_sequential()
    assert a.digest() != b.digest()
    ids, words = b.final_states(100)
    assert ids.tolist() == list(range(100))
    assert int(words[:, 0].sum()) > 0
