#!/bin/bash
# Builds libspliser_b200.so with AddressSanitizer + UBSan in a scratch copy of the tracked files, runs the CPU test
# suite and the mutation fuzzer (tests/tools/fuzz_native.py) against it, and prints whatever the sanitizers reported.
#   tests/tools/asan_run.sh [seconds-per-fuzz-target] [seed]
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
WORK=${SPLISER_ASAN_DIR:-/tmp/spliser_asan}
rm -rf "$WORK" && mkdir -p "$WORK"
(cd "$ROOT" && git ls-files -co --exclude-standard | tar cf - -T - 2>/dev/null) | tar xf - -C "$WORK"
cd "$WORK"
sed -i 's/"-O3"/"-O1"/' spliser_b200/build.py
SPLISER_NVCC_FLAGS="-Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer,-g" python -m spliser_b200.build --force >/dev/null
make -s -C oracle
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:abort_on_error=1:log_path=$WORK/asan.log
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0:log_path=$WORK/ubsan.log
export LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)
python -m pytest tests/ -x -q -m "not gpu" -p no:cacheprovider
python tests/tools/fuzz_native.py "${1:-10}" "${2:-1}"
unset LD_PRELOAD
if ls "$WORK"/asan.log* "$WORK"/ubsan.log* >/dev/null 2>&1; then
    echo "SANITIZER REPORTS:"; head -n 60 "$WORK"/asan.log* "$WORK"/ubsan.log* 2>/dev/null; exit 1
fi
echo "sanitizers: clean"
