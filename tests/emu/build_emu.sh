#!/bin/sh
# Builds the SIMT-emulated test library (TEST INFRASTRUCTURE ONLY; see emu.h).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
g++ -O2 -g -std=c++17 -fPIC -shared -DCS_EMU -x c++ -I"$here" -I"$root/causalstump_b200/csrc" \
    "$root/causalstump_b200/csrc/api.cu" -x c++ "$here/emu.cpp" -o "$here/libcs_emu.so"
