#!/bin/bash
# Build a kernel variant of libnnsp_b200.so next to the product library: bash tools/build_variant.sh NAME "-DFLAG=.. ..."
# -> build/variants/libnnsp_b200_NAME.so (select it with NNSP_B200_LIB=...; measurements only, never shipped).
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/build/variants"
make -s -j8 -C "$root/nnsp_b200/csrc" OUT="$root/build/variants/libnnsp_b200_$name.so" OBJ="$root/build/obj_$name" EXTRA_NVFLAGS="$flags"
echo "$root/build/variants/libnnsp_b200_$name.so"
