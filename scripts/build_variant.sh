#!/bin/bash
# scripts/build_variant.sh NAME [GIT_REV|-] [extra nvcc flags...]: build breeze.jl_b200/csrc/variants/NAME.so from the working tree
# (GIT_REV = "-") or from a committed revision, with extra -D flags, for scripts/variant_bench.py.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; REV=${2:--}; shift; shift || true
OUT=$ROOT/breeze.jl_b200/csrc/variants; mkdir -p $OUT
SRC=$ROOT/breeze.jl_b200/csrc
if [ "$REV" != "-" ]; then
  TMP=$(mktemp -d); mkdir -p $TMP/breeze.jl_b200 $TMP/include
  git -C $ROOT archive $REV breeze.jl_b200/csrc include | tar -x -C $TMP
  SRC=$TMP/breeze.jl_b200/csrc
fi
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -O2 "$@" $SRC/api.cu -o $OUT/$NAME.so -ldl
echo built $OUT/$NAME.so
