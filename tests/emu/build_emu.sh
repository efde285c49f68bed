#!/bin/sh
# Builds tests/emu/_build/libtb200_emu.so: the product's kernel sources compiled for the host
# (-DTB200_HOST_EMU).  Test infrastructure only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../tiberate_fhe_b200/csrc"
mkdir -p "$HERE/_build"
CXX=/usr/bin/g++
[ -x "$CXX" ] || CXX=g++
$CXX -std=c++20 -O2 -g -ffp-contract=off -fPIC -shared -pthread -DTB200_HOST_EMU -I"$SRC" \
  -Wno-unknown-pragmas -Wno-attributes \
  -x c++ "$SRC/tb200.cu" -x c++ "$HERE/emu_runtime.cpp" -o "$HERE/_build/libtb200_emu.so"
echo "$HERE/_build/libtb200_emu.so"
