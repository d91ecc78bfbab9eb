#!/bin/sh
# Developer tool: compiles the engine + kernels for the HOST (sequential emulation, see cuda_emu.h) into
# tools/emu/libssb_emu.so. Not part of the product, not built by __graft_entry__.build().
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
g++ -std=c++17 -O1 -g -DSSB_EMU -fPIC -shared -Wno-unknown-pragmas -I "$HERE" \
    -x c++ "$ROOT/scalesim_b200/csrc/ssb_engine.cu" -x c++ "$ROOT/scalesim_b200/csrc/phold_host.cc" \
    -o "$HERE/libssb_emu.so" -ldl
echo "built $HERE/libssb_emu.so"
