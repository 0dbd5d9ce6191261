#!/usr/bin/env bash
# Builds the UNMODIFIED reference (read-only /root/reference) into baseline/_ref/dgl for
# `bench.py --impl reference`.  See DESIGN.md section 4 for why the pip route does not work.
# Usage: script/build_reference.sh [jobs]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
JOBS="${1:-$(nproc)}"
SRC=/tmp/dgl_src
BLD=/tmp/dgl_ref_build
DST="$ROOT/baseline/_ref"

if [ ! -d "$SRC" ]; then
  cp -r /root/reference "$SRC"
fi
mkdir -p "$BLD"
cd "$BLD"
export CC=/usr/bin/gcc CXX=/usr/bin/g++ CUDAHOSTCXX=/usr/bin/g++
cmake "$SRC" -DBUILD_TYPE=release -DUSE_CUDA=ON -DCUDA_ARCH_NAME=Manual -DCUDA_ARCH_BIN=100 \
  -DCUDA_ARCH_PTX=100 -DBUILD_GRAPHBOLT=OFF -DBUILD_SPARSE=OFF -DUSE_LIBXSMM=OFF \
  -DBUILD_CPP_TEST=OFF -DBUILD_TORCH=ON -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc \
  -DPYTHON_EXECUTABLE="$(command -v python)" > cmake.log 2>&1
make -j"$JOBS" > make.log 2>&1

rm -rf "$DST/dgl"
mkdir -p "$DST"
cp -r "$SRC/python/dgl" "$DST/dgl"
cp "$BLD/libdgl.so" "$DST/dgl/"
mkdir -p "$DST/dgl/tensoradapter/pytorch"
find "$BLD" -name 'libtensoradapter_pytorch_*.so' -exec cp {} "$DST/dgl/tensoradapter/pytorch/" \;
echo "reference installed into $DST/dgl"
PYTHONPATH="$DST" python -c "import dgl; print('dgl', dgl.__version__)"
