"""CPU logic checks on the emulated kernels (tools/emu) where there is no GPU. First the MULTI-PARTITION path
(SURVEY.md §8e): the kernel sources compiled for
the host (tools/emu, a developer tool the product cannot load) with one emulated engine per partition, each on its own
host thread, mailboxes exchanged by address — cross-partition events and anti-messages through k_route / the receive
pass, GVT through k_gvt, sequence stamps. The union of the partitions' committed multisets must equal the oracle's for
every world size (the invariant of §8e). It complements, and does not replace, the `-m gpu` tests on two B200s
(tests/test_gpu_multi.py): races and NVLink visibility exist only there. Runs in a subprocess because the emulation is
selected per process."""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu_results():
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    subprocess.run(["sh", os.path.join(ROOT, "tools", "emu", "build.sh")], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                   timeout=600)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu", "check_multi.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_phold_partitions_commit_the_oracles_multiset(emu_results):
    for k in ("phold_w2_win1", "phold_w2_win3", "phold_w3_win2", "phold_w4_win1"):
        r = emu_results[k]
        assert r["equal"] and r["committed"] == 72351 and r["remote"] > 0, k
    assert emu_results["phold_w2_win1"]["rollbacks"] == 0
    assert emu_results["phold_w2_win3"]["rollbacks"] > 0 and emu_results["phold_w2_win3"]["antis"] > 0


def test_state_mutating_model_across_partitions(emu_results):
    r = emu_results["stateful_w3"]
    assert r["equal"] and r["states_equal"] and r["rollbacks"] > 0


def test_traffic_stripes_commit_the_oracles_multiset(emu_results):
    for k in ("grid_w2_win1", "grid_w4_win3"):
        assert emu_results[k]["equal"] and emu_results[k]["remote"] > 0, k
    assert emu_results["grid_w4_win3"]["rollbacks"] > 0


def test_several_windows_per_synchronisation(emu_results):
    """ssb_set_sync_interval on the interpreted graph loop: fewer exchanges than rounds, stragglers across partitions
    repaired by rollbacks, same committed multiset."""
    for k in ("grid_w2_sync2", "grid_w4_sync8", "phold_w3_sync4"):
        r = emu_results[k]
        assert r["equal"] and r["exchanges"] < r["rounds"] and r["rollbacks"] > 0, k
    assert emu_results["phold_w2_stream"]["equal"]


def test_partitions_started_from_one_scenario_image(emu_results):
    r = emu_results["grid_w4_from_image"]
    assert r["equal"] and r["remote"] > 0


def test_nccl_exchange_mode(emu_results):
    """the ncclSend / ncclRecv + all-reduce exchange (ssb_comm_init; --exchange nccl) over the emulation's in-process NCCL:
    counts all-to-all, payload all-to-all, per-peer receive pass, GVT by all-reduce"""
    for k in ("phold_w3_nccl", "grid_w4_nccl"):
        assert emu_results[k]["equal"] and emu_results[k]["remote"] > 0 and emu_results[k]["rollbacks"] > 0, k


def test_device_parity_tests_pass_on_the_emulated_kernels(emu_results):
    """The `-m gpu` files that fit a CPU budget — device KATs of logical_process_test.cc, TrafficSim, scenario images,
    overflow behaviour, the drop-in binaries — run against the emulated kernels in a subprocess: a logic regression net
    where there is no GPU (the same tests run on the B200 with `-m gpu`; only those runs are parity claims)."""
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tools", "emu"), SCALESIM_B200_APPS_DIR=os.path.join(ROOT, "tools", "emu", "apps"))
    files = [os.path.join(ROOT, "tests", f) for f in ("test_gpu_kat.py", "test_gpu_scenario.py", "test_gpu_zz_overflow.py",
                                                       "test_gpu_dropin.py", "test_gpu_traffic.py")]
    p = subprocess.run([sys.executable, "-m", "pytest", "-p", "emu_plugin", "-m", "gpu", "-q", "-x", "-k", "not full_size",
                        "-p", "no:cacheprovider"] + files, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                       cwd=ROOT, timeout=1200)
    tail = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
    assert p.returncode == 0 and " passed" in tail and "failed" not in tail, p.stdout[-3000:]
