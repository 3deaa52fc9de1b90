#!/bin/bash
# Sanitizer passes over the kernels' real source on the CPU (tests/emu): the host emulation runs every
# lane as a host thread that only synchronises where the kernel does, so AddressSanitizer sees every
# global / __shared__ (static) access and ThreadSanitizer sees every pair of unsynchronised accesses.
#   scripts/cpu_sanitize.sh asan   -> -fsanitize=address,undefined over all emulation tests
#   scripts/cpu_sanitize.sh tsan   -> -fsanitize=thread over the walk / tree / direct emulation tests
set -e
cd "$(dirname "$0")/.."
mode=${1:-asan}
bak=$(mktemp -d)
cp tests/emu/lib*_emu.so "$bak"/ 2>/dev/null || true
restore() { cp "$bak"/lib*_emu.so tests/emu/ 2>/dev/null || true; touch tests/emu/lib*_emu.so 2>/dev/null || true; }
trap restore EXIT
if [ "$mode" = asan ]; then flags="-fsanitize=address,undefined -fno-omit-frame-pointer"; units="walk tree direct ic"
else flags="-fsanitize=thread"; units="walk tree direct"; fi
for f in $units; do
  g++ -O1 -g -std=c++20 -shared -fPIC -pthread -DGH_EMU_THREADS $flags -I"${CUDA_HOME:-/usr/local/cuda}/include" -o tests/emu/lib${f}_emu.so tests/emu/${f}_emu.cpp
done
touch tests/emu/lib*_emu.so
tests=""; for f in $units; do tests="$tests tests/test_${f}_emu.py"; done
if [ "$mode" = asan ]; then
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    python -m pytest $tests -q -p no:cacheprovider
else
  TSAN_OPTIONS="halt_on_error=0:report_signal_unsafe=0:exitcode=0" LD_PRELOAD="$(gcc -print-file-name=libtsan.so)" \
    python -m pytest $tests -q -p no:cacheprovider 2>&1 | tee /tmp/gh_tsan.log | tail -3
  echo "ThreadSanitizer warnings: $(grep -c 'WARNING: ThreadSanitizer' /tmp/gh_tsan.log || true)"
fi
