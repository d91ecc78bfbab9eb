"""The drop-in binaries: the reference's UNMODIFIED application sources (phold.hpp, traffic_sim.hpp, both main_*.cc)
compiled against include/scalesim and linked to libscalesim_b200.so. Their stdout must be the reference's."""
import os
import subprocess
import tempfile

import pytest

from fixtures import golden, golden_lines, sha256_lines, write_ring_files

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APPS = os.environ.get("SCALESIM_B200_APPS_DIR", os.path.join(ROOT, "apps"))      # (tools/emu builds its own copies)


def _run(exe, args, env=None):
    path = os.path.join(APPS, exe)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs the reference tree at build time)")
    p = subprocess.run([path] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, **(env or {})), timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("RE,")], p.stderr


def test_phold_dropin_equals_reference():
    lines, err = _run("PHOLD", [100, 1600, 0.1, 1.0, 1, 50], {"SCALESIM_B200_VERIFY": "400"})
    assert sorted(lines) == golden_lines("phold_100_1600")
    assert "# of output events             : 7517" in err
    assert "0 mismatches" in err           # App::event_handler on the host agrees with the device twin


def test_phold_dropin_optimistic_window():
    lines, err = _run("PHOLD", [1000, 4000, 0.25, 2.0, 1, 50], {"SCALESIM_B200_WINDOW": "4"})
    assert sha256_lines(lines) == golden("phold_1000_4000")["sha256"]


def test_trafficsim_dropin_equals_reference():
    g = golden("ring")
    with tempfile.TemporaryDirectory() as td:
        rd, trip, part = write_ring_files(td)
        part1 = os.path.join(td, "part1")
        with open(part1, "w") as f:
            f.write("0\n" * 10)
        lines, err = _run("TrafficSim", [part1, rd, trip], {"SCALESIM_B200_VERIFY": "3"})
    assert sorted(lines) == g["lines"]
    assert "0 mismatches" in err


def test_trafficsim_dropin_diff_init_and_repeat_equal_reference(tmp_path):
    """TrafficSim --diff_init on the ring, then --diff_repeat with traffic/ring/add_ev.csv:
    38 lines and the reference's store records, then the reference's 19 lines."""
    from scalesim_b200 import store
    g = golden("ring_diff")
    rd, trip, part = write_ring_files(str(tmp_path))
    part1 = str(tmp_path / "part1")
    with open(part1, "w") as f:
        f.write("0\n" * 10)
    ae = str(tmp_path / "add_ev.csv")
    with open(ae, "w") as f:
        f.write("".join(l + "\n" for l in g["ae_lines"]))
    flags = ["--sim_id=999", f"--store_dir={tmp_path}/st"]
    lines, err = _run("TrafficSim", flags + ["--diff_init", part1, rd, trip])
    assert sorted(lines) == g["init_lines"]
    hs = store.HistoryStore.load(f"{tmp_path}/st", 999, 0)
    assert hs.event_keys().tolist() == sorted(g["store"]["event"])
    assert hs.cancel_keys().tolist() == sorted(g["store"]["cancel"])
    assert hs.state_keys().tolist() == sorted(g["store"]["state"])
    lines, err = _run("TrafficSim", flags + ["--diff_repeat", part1, rd, ae])
    assert sorted(lines) == g["repeat_union"]
    assert "# of output events             : 19" in err
    de = str(tmp_path / "delete_ev.csv")           # traffic/ring/delete_ev.csv: one RE line
    with open(de, "w") as f:
        f.write("".join(l + "\n" for l in g["de_lines"]))
    lines, err = _run("TrafficSim", flags + ["--diff_repeat", part1, rd, de])
    assert sorted(lines) == g["delete_union"]
    scq = str(tmp_path / "change_state.csv")       # traffic/ring/change_state.csv: the SC query
    with open(scq, "w") as f:
        f.write("".join(l + "\n" for l in g["sc_lines"]))
    lines, err = _run("TrafficSim", flags + ["--diff_repeat", part1, rd, scq])
    from test_gpu_diff import ring_changed_network_lines
    init, full = set(g["init_lines"]), set(ring_changed_network_lines())
    new = set(l for l in lines if l not in init)
    assert sorted((init - (init - full)) | new) == sorted(full)                          # init + repeat == one run on the changed network
    assert new == set(g["sc_new_lines"]) & full                                          # the reference's (right) new lines
    assert all(set(lines) <= set(v["lines"]) for v in g["sc_variants"])                 # ... and nothing it does not print


def test_phold_dropin_diff_init_and_repeat_equal_reference(tmp_path):
    g = golden("phold_diff")
    args = [str(x) for x in g["args"]] + [str(g["what_ifs"]), "50"]
    flags = ["--sim_id=3", f"--store_dir={tmp_path}/st"]
    lines, err = _run("PHOLD", flags + ["--diff_init"] + args)
    assert sorted(lines) == g["init_lines"]
    lines, err = _run("PHOLD", flags + ["--diff_repeat"] + args)
    assert sorted(lines) == g["repeat_union"]


def test_dropin_repeat_without_store_fails_loudly(tmp_path):
    path = os.path.join(APPS, "PHOLD")
    if not os.path.exists(path):
        pytest.skip("not built")
    p = subprocess.run([path, "--diff_repeat", "--sim_id=1", f"--store_dir={tmp_path}/none", "100", "1600", "0.1", "1.0", "1", "50"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 1 and "no history store" in p.stderr


def test_phold_dropin_writes_and_starts_from_a_scenario_image(tmp_path):
    """SCALESIM_B200_SCENARIO_OUT keeps what phold.hpp's generators produced as a scenario image; a second run with
    SCALESIM_B200_SCENARIO starts from it (no init_events, no per-event objects) and prints the same reference lines."""
    img = str(tmp_path / "phold.ssbscn")
    args = [100, 1600, 0.1, 1.0, 1, 50]
    lines, err = _run("PHOLD", args, {"SCALESIM_B200_SCENARIO_OUT": img})
    assert sorted(lines) == golden_lines("phold_100_1600") and "Scenario image written" in err
    lines, err = _run("PHOLD", args, {"SCALESIM_B200_SCENARIO": img, "SCALESIM_B200_WINDOW": "3"})
    assert sorted(lines) == golden_lines("phold_100_1600")
    assert "Total events num (this partition): 1600" in err and "Total states num (this partition): 100" in err
    # an image of another application is refused
    p = subprocess.run([os.path.join(APPS, "TrafficSim"), "x", "y", "z"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, SCALESIM_B200_SCENARIO=img), timeout=120)
    assert p.returncode == 1 and "another application" in p.stderr


def test_trafficsim_dropin_from_a_compiled_csv_image(tmp_path):
    """The ring's CSV files compiled natively (ssb_traffic_compile) -> TrafficSim started from the image: the
    reference's 38 lines; --diff_init from the image writes the store the reference writes, and --diff_repeat on it
    prints the reference's 19 lines."""
    from scalesim_b200 import scenario, store
    g, gd = golden("ring"), golden("ring_diff")
    rd, trip, part = write_ring_files(str(tmp_path))
    part1 = str(tmp_path / "part1")
    with open(part1, "w") as f:
        f.write("0\n" * 10)
    img = str(tmp_path / "ring.ssbscn")
    scenario.compile_traffic_csv(rd, trip, img, partition_file=part1)
    env = {"SCALESIM_B200_SCENARIO": img}
    lines, err = _run("TrafficSim", ["-", "-", "-"], env)              # the file arguments are not read
    assert sorted(lines) == g["lines"]
    flags = ["--sim_id=5", f"--store_dir={tmp_path}/st"]
    lines, err = _run("TrafficSim", flags + ["--diff_init", "-", "-", "-"], env)
    assert sorted(lines) == gd["init_lines"]
    hs = store.HistoryStore.load(f"{tmp_path}/st", 5, 0)
    assert hs.event_keys().tolist() == sorted(gd["store"]["event"])
    assert hs.cancel_keys().tolist() == sorted(gd["store"]["cancel"])
    assert hs.state_keys().tolist() == sorted(gd["store"]["state"])
    ae = str(tmp_path / "add_ev.csv")
    with open(ae, "w") as f:
        f.write("".join(l + "\n" for l in gd["ae_lines"]))
    lines, err = _run("TrafficSim", flags + ["--diff_repeat", part1, rd, ae])
    assert sorted(lines) == gd["repeat_union"]
