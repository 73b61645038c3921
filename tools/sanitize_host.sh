#!/bin/bash
# The CPU test suite against an AddressSanitizer + UndefinedBehaviorSanitizer build of libvfkio.so (BGZF / BAM decoder, inflate,
# INFO text).  The release library is put back afterwards.   usage: tools/sanitize_host.sh
set -u
cd "$(dirname "$0")/.."
lib=vafator_b200/libvfkio.so
keep=$(mktemp)
cp "$lib" "$keep"
trap 'cp "$keep" "$lib"; touch "$lib"; rm -f "$keep"' EXIT
g++ -O1 -g -std=c++17 -fPIC -shared -fopenmp -fsanitize=address,undefined -fno-omit-frame-pointer -I include \
    -o "$lib" $(ls vafator_b200/csrc/vfkio_*.cpp | grep -v vfkio_synth.cpp) -lz -lpthread || exit 1
touch "$lib"
log=$(mktemp)
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
    python -m pytest tests -q -m "not gpu" -p no:cacheprovider 2>&1 | tee "$log" | tail -3
n=$(grep -c "runtime error\|AddressSanitizer" "$log")
echo "sanitizer reports: $n"
rm -f "$log"
exit "$n"
