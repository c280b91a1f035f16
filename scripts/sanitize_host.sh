#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer (or ThreadSanitizer) run of the C++ host runtime (csrc_host/: record /
# LMDB / LevelDB readers, snappy decoder, batch-loader thread pool, LibSVM parser) under the CPU tests that drive it.
# The parsers consume files from outside the framework, so malformed input must fail with an exception, never with an
# out-of-bounds access.  (the reference has no sanitizer targets.)
#   scripts/sanitize_host.sh [asan|tsan]
# The instrumented module is built into a scratch directory and selected with POSEIDON_HOST_SO; the in-tree .so is untouched.
set -e
cd "$(dirname "$0")/.."
MODE=${1:-asan}
OUT=$(mktemp -d)/poseidon_b200_host.so
PYINC=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
PBINC=$(python -c "import pybind11; print(pybind11.get_include())")
if [ "$MODE" == "tsan" ]; then
  SAN="-fsanitize=thread"; RT=$(gcc -print-file-name=libtsan.so)
else
  SAN="-fsanitize=address,undefined -fno-sanitize-recover=undefined"; RT=$(gcc -print-file-name=libasan.so)
fi
g++ -O1 -g -fno-omit-frame-pointer $SAN -std=c++17 -shared -fPIC -pthread -I "$PBINC" -I "$PYINC" -I csrc_host \
    csrc_host/record_loader.cpp csrc_host/leveldb_reader.cpp csrc_host/libsvm_parser.cpp -o "$OUT"
echo "built $OUT ($MODE)"
# leak checking is off: the interpreter itself never frees its arenas
# libstdc++ must be loaded together with the sanitizer runtime (python itself does not link it), or the __cxa_throw
# interceptor has nothing to forward to
STD=$(gcc -print-file-name=libstdc++.so.6)
POSEIDON_FUZZ_ITERS=${POSEIDON_FUZZ_ITERS:-200} POSEIDON_HOST_SO="$OUT" LD_PRELOAD="$RT $STD" ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
  TSAN_OPTIONS="report_signal_unsafe=0" \
  python -m pytest tests/test_native_loader.py tests/test_lmdb_reader.py tests/test_leveldb_reader.py tests/test_ml_helpers.py tests/test_host_fuzz.py tests/test_leveldb_writer.py \
    -q -x -p no:cacheprovider 2>&1 | grep -v 'python3.12+0x' | tail -25
