#!/bin/bash
# Profiling / ablation build: recompile ONE csrc file with extra defines and link it with the regular objects into
# linnaeus_b200/variants/<name>.so (git-ignored).  Use with LNX_LIB_PATH=... python tools/prof_mlp_fused.py ...
#   tools/build_variant.sh lnx_mlp_fused.cu dbg48 -DLNX_DBG=48
set -e
cd "$(dirname "$0")/.."
src=$1; name=$2; shift 2
python -m linnaeus_b200._build >/dev/null 2>&1
mkdir -p linnaeus_b200/variants
obj=linnaeus_b200/variants/${name}.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr "$@" \
  -Iinclude -c linnaeus_b200/csrc/$src -o $obj
others=$(ls linnaeus_b200/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o linnaeus_b200/variants/${name}.so $obj $others -gencode arch=compute_100a,code=sm_100a -lcudart
echo linnaeus_b200/variants/${name}.so
