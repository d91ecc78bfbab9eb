#!/bin/sh
# Developer tool: compiles the engine + kernels for the HOST (sequential emulation, see cuda_emu.h) into
# tools/emu/libssb_emu.so. Not part of the product, not built by __graft_entry__.build().
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
g++ -std=c++17 -O1 -g -DSSB_EMU -fPIC -shared -Wno-unknown-pragmas -I "$HERE" \
    -x c++ "$ROOT/scalesim_b200/csrc/ssb_engine.cu" -x c++ "$ROOT/scalesim_b200/csrc/phold_host.cc" \
    -x c++ "$ROOT/scalesim_b200/csrc/scenario_host.cc" \
    -o "$HERE/libssb_emu.so" -ldl
echo "built $HERE/libssb_emu.so"
# the drop-in applications against the emulation (only where the reference tree exists): tools/emu/apps/{PHOLD,TrafficSim};
# SCALESIM_B200_APPS_DIR=tools/emu/apps points tests/test_gpu_dropin.py at them
if [ -d /root/reference ]; then
  mkdir -p "$HERE/apps"
  for app in phold:PHOLD trafficsim:TrafficSim; do
    src=${app%%:*}; bin=${app##*:}
    g++ -std=c++17 -O1 -g -w -pthread -I "$ROOT/include" -I "$ROOT/include/compat" -I /root/reference/src \
        "$ROOT/apps/${src}_dropin.cc" -o "$HERE/apps/$bin" -L "$HERE" -lssb_emu -Wl,-rpath,"$HERE"
  done
  echo "built $HERE/apps/PHOLD $HERE/apps/TrafficSim"
fi
