#!/usr/bin/env bash
# Build the two in-tree native libraries (dgl_b200/lib/libdglb200_{cpu,cuda}.so).
#   script/build.sh            incremental (object files are hashed on source + flags + headers)
#   script/build.sh -f         force a full rebuild
#   script/build.sh -t         build, then run the CPU test-suite and the native C++ unit tests
# Environment: CUDA_HOME (default /usr/local/cuda), DGLB200_CXX (host compiler), DGLB200_BUILD_JOBS.
# The CUDA library is always compiled for sm_100a (-gencode arch=compute_100a,code=sm_100a -lineinfo); there is no
# multi-arch switch on purpose. Counterpart of the reference's script/build_dgl.sh.
set -euo pipefail
cd "$(dirname "$0")/.."
force=0; tests=0
while getopts "ft" o; do case "$o" in f) force=1;; t) tests=1;; *) echo "usage: $0 [-f] [-t]"; exit 2;; esac; done
[ "$force" = 1 ] && rm -rf dgl_b200/lib/obj dgl_b200/lib/*.so
python -c "import __graft_entry__ as g; g.build()"
if [ "$tests" = 1 ]; then
  python -m pytest tests -x -q -m "not gpu"
fi
