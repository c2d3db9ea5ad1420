#!/bin/bash
# Build an experimental variant of the library next to the product one (tuning helper):
#   tools/build_variant.sh <name> "<extra nvcc flags for kernels_pair.cu>" ["<extra flags for kernels.cu>"]
# -> pairing_b200/lib/exp_<name>.so (git-ignored; select it with PAIRING_B200_LIB=...).
set -e
cd "$(dirname "$0")/../pairing_b200/csrc"
name=$1; pflags=$2; kflags=$3
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default"
mkdir -p _obj/$name ../lib
$NVCC $FLAGS $pflags -c -o _obj/$name/kernels_pair.o kernels_pair.cu &
if [ -n "$kflags" ] || [ ! -f _obj/kernels.o ]; then
  $NVCC $FLAGS $kflags -c -o _obj/$name/kernels.o kernels.cu
  KOBJ=_obj/$name/kernels.o
else
  KOBJ=_obj/kernels.o
fi
wait
[ -f _obj/kernels_wide.o ] || $NVCC $FLAGS -c -o _obj/kernels_wide.o kernels_wide.cu
[ -f _obj/mgpu.o ] || $NVCC $FLAGS -c -o _obj/mgpu.o mgpu.cu
# the multi-pairing unit takes the same extra flags as the pairing unit
$NVCC $FLAGS $pflags -c -o _obj/$name/kernels_mm.o kernels_mm.cu
$NVCC $ARCH -shared -o ../lib/exp_$name.so $KOBJ _obj/$name/kernels_pair.o _obj/$name/kernels_mm.o _obj/kernels_wide.o _obj/mgpu.o -lpthread
echo "built pairing_b200/lib/exp_$name.so"
cuobjdump --dump-resource-usage ../lib/exp_$name.so 2>/dev/null | grep -A1 "k_pair_millerILb1\|k_pair_multi_millerPK" | grep -o "Function [^:]*\|REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*" | paste - - - -
