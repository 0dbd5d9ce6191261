#!/usr/bin/env bash
# Memory-safety / data-race gate for the host native code (csrc/cpu). The reference ships no sanitizer configuration
# (SURVEY 5.2).
#   script/run_sanitizers.sh             C++ unit tests under AddressSanitizer + UndefinedBehaviorSanitizer
#   script/run_sanitizers.sh --tsan      C++ unit tests under ThreadSanitizer. Meaningful for the pthread code (RPC
#                                        transport, caches): libgomp is not instrumented, so OpenMP regions report
#                                        fork/join false positives — those kernels are checked for races by
#                                        tests/test_thread_invariance.py (bit-identical results for 1 / 3 / 8 threads)
#   script/run_sanitizers.sh --python [pytest args]
#                                        build libdglb200_cpu.so with ASan+UBSan into a scratch directory and run the
#                                        CPU pytest suite against it (DGLB200_LIBDIR), with libasan and libstdc++
#                                        preloaded so C++ exceptions keep working under the interceptors
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${TMPDIR:-/tmp}/dglb200_san"
mkdir -p "$OUT/lib"
CXX=${DGLB200_CXX:-$([ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)}
COMMON="-O1 -g -std=c++17 -fopenmp -pthread -fno-omit-frame-pointer -I$ROOT/dgl_b200/csrc"
export OMP_NUM_THREADS=${OMP_NUM_THREADS:-4}      # several threads so that races / per-thread scratch bugs can show
export ASAN_OPTIONS=${ASAN_OPTIONS:-detect_leaks=0:abort_on_error=1}
export UBSAN_OPTIONS=${UBSAN_OPTIONS:-print_stacktrace=1}
export TSAN_OPTIONS=${TSAN_OPTIONS:-ignore_noninstrumented_modules=1}   # libgomp itself is not instrumented

case "${1:-}" in
  --python)
    shift
    $CXX $COMMON -fPIC -shared -fsanitize=address,undefined "$ROOT"/dgl_b200/csrc/cpu/*.cc -o "$OUT/lib/libdglb200_cpu.so" -lrt
    export DGLB200_LIBDIR="$OUT/lib"
    export LD_PRELOAD="$($CXX -print-file-name=libasan.so) $($CXX -print-file-name=libstdc++.so.6)"
    cd "$ROOT"
    # the multi-process server test spawns fresh interpreters that would each re-run the preload dance: skipped
    exec python -m pytest tests -q -m "not gpu" -p no:cacheprovider \
      --deselect tests/test_distributed.py::test_multiprocess_servers "$@"
    ;;
  --tsan)
    $CXX $COMMON -fsanitize=thread "$ROOT"/dgl_b200/csrc/cpu/*.cc "$ROOT/tests/cpp/test_native.cc" -o "$OUT/test_native_tsan" -lrt
    exec "$OUT/test_native_tsan"
    ;;
  *)
    $CXX $COMMON -fsanitize=address,undefined -fno-sanitize-recover=undefined "$ROOT"/dgl_b200/csrc/cpu/*.cc \
      "$ROOT/tests/cpp/test_native.cc" -o "$OUT/test_native_asan" -lrt
    exec "$OUT/test_native_asan"
    ;;
esac
