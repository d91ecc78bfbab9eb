"""pytest plugin (developer tool): run the `-m gpu` parity tests against the sequential HOST emulation of the
kernels (tools/emu/libssb_emu.so) to debug their logic on a machine without a GPU:

    tools/emu/build.sh && PYTHONPATH=tools/emu python -m pytest -p emu_plugin tests/test_gpu_phold.py -m gpu -k "not 1m"

Nothing in the product can select this library; the plugin swaps the path inside the test process only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import scalesim_b200.capi as capi  # noqa: E402

# EMU_LIB: another build of the emulation, e.g. one with -fsanitize=address (run python under LD_PRELOAD=libasan.so)
_EMU = os.environ.get("EMU_LIB") or os.path.join(ROOT, "tools", "emu", "libssb_emu.so")
capi.library_path = lambda: _EMU


def pytest_configure(config):
    import __graft_entry__ as g
    g.build_library = lambda force=False: _EMU      # do not rebuild the device library for an emulation run
