#!/bin/sh
# AddressSanitizer + ThreadSanitizer runs of the host-compiled kernel sources (see sanitize_case.py).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../.."
SRC="$ROOT/tiberate_fhe_b200/csrc"
OUT="${1:-$ROOT/profiles}"
mkdir -p "$HERE/_build" "$OUT"
for san in address thread; do
  g++ -std=c++20 -O1 -g -ffp-contract=off -fPIC -shared -pthread -fsanitize=$san -fno-omit-frame-pointer \
    -DTB200_HOST_EMU -I"$SRC" -Wno-unknown-pragmas -Wno-attributes \
    -x c++ "$SRC/tb200.cu" -x c++ "$HERE/emu_runtime.cpp" -o "$HERE/_build/libtb200_emu_$san.so"
  rt=$(gcc -print-file-name=lib$( [ $san = address ] && echo asan || echo tsan ).so)
  log="$OUT/r02_emu_${san}_sanitizer.log"
  echo "== $san: LD_PRELOAD=$rt python tests/emu/sanitize_case.py" > "$log"
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" \
    LD_PRELOAD="$rt" python "$HERE/sanitize_case.py" "$HERE/_build/libtb200_emu_$san.so" >> "$log" 2>&1 || echo "exit code $?" >> "$log"
  echo "reports: $(grep -c -E 'ERROR: AddressSanitizer|WARNING: ThreadSanitizer' "$log" || true)" >> "$log"
  tail -4 "$log"
done
