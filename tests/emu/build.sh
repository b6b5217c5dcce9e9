#!/bin/sh
# TEST TOOL: host instantiation of embree_b200/csrc/rt_core.cuh (see emu.cpp header comment)
set -e
cd "$(dirname "$0")"
mkdir -p _build
g++ -O2 -g -std=c++17 -fPIC -shared -mfma -ffp-contract=off -o _build/libemu.so emu.cpp
