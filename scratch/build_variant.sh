#!/bin/bash
# usage: build_variant.sh NAME [nvcc -D flags...]   (ROWS=path overrides pd_warp_rows.cuh)
set -e
cd /root/repo
NAME=$1; shift
TMP=$(mktemp -d)
cp planedepth_b200/csrc/*.cu planedepth_b200/csrc/*.cuh $TMP/
mkdir -p $TMP/../../include 2>/dev/null || true
if [ -n "$ROWS" ]; then cp $ROWS $TMP/pd_warp_rows.cuh; fi
sed -i 's#"../../include/planedepth_b200.h"#"planedepth_b200.h"#' $TMP/pd_device.cuh
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include "$@" -o scratch/variants/lib_$NAME.so $TMP/pd_abi.cu
rm -rf $TMP
echo built $NAME
