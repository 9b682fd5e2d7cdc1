#!/bin/bash
# build_ref_cuda.sh <radius literal> <output binary> [errorcheck 1|0] -- the reference's own CUDA build, headless (see ref_headless_main.cpp).
# errorcheck 0 = cudaUtil.h:8 "#define ERRORCHECK" set to 0: the device-wide sync the reference issues after every launch
# (cudaUtil.h:13-16) is compiled out, for a kernel-only timing next to the as-shipped one.
# Compiles the reference's sources where they lie under $REF; the only modified file, restir.cu, is patched into a
# mktemp directory (scope braces for nvcc, unused GI kernel #if 0'd, radius literal overridable) and
# removed again (the history-reservoir pointer is also made extern so that the driver can dump it).  TEST / BASELINE INFRASTRUCTURE ONLY.
set -e
RADIUS=$1; OUT=$2; ERRCHK=${3:-1}
REF=${REF:-/root/reference}; CXX=/usr/bin/g++; NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
HERE=$(cd "$(dirname "$0")" && pwd)
OBJ=$(mktemp -d)
trap 'rm -rf "$OBJ"' EXIT
HOSTFLAGS="-std=c++17 -O2 -w -D__stdcall= -include math.h -include string.h -include float.h -I$REF/src -I$REF/external/include -I/usr/local/cuda/include -I$HERE"
CUFLAGS="-std=c++17 -O3 -w -gencode arch=compute_100a,code=sm_100a -D__stdcall= -ccbin $CXX -I$REF/src -I$REF/external/include -I$HERE"
for f in scene bvh mathUtil utilities image common stb tiny_obj_loader; do $CXX $HOSTFLAGS -c $REF/src/$f.cpp -o $OBJ/$f.o; done
# the two TUs that launch kernels are compiled from mktemp copies next to a copy of cudaUtil.h carrying the ERRORCHECK value
# (a quoted #include looks beside the including file first); the copies are deleted right after nvcc ran
sed -e "s/^#define ERRORCHECK 1/#define ERRORCHECK $ERRCHK/" $REF/src/cudaUtil.h > $OBJ/cudaUtil.h
cp $REF/src/gbuffer.cu $OBJ/gbuffer.cu
$NVCC $CUFLAGS -c $OBJ/gbuffer.cu -o $OBJ/gbuffer.o
rm -f $OBJ/gbuffer.cu
$NVCC $CUFLAGS -c $REF/src/denoiser.cu -o $OBJ/denoiser.o
sed -e '140i {' -e '228i }' -e '233i #if 0' -e '417i #endif' -e '448i #if 0' -e '477i #endif' \
    -e "s/const float Radius = 5.f;/const float Radius = $RADIUS;/" \
    -e 's/^static DirectReservoir\* devLastDirectReservoir/DirectReservoir* devLastDirectReservoir/' $REF/src/restir.cu > $OBJ/restir.cu
$NVCC $CUFLAGS -c $OBJ/restir.cu -o $OBJ/restir.o
rm -f $OBJ/restir.cu $OBJ/cudaUtil.h
$NVCC $CUFLAGS -x cu -c $HERE/ref_headless_main.cpp -o $OBJ/main.o
mkdir -p "$(dirname "$OUT")"
$NVCC -ccbin $CXX -o "$OUT" $OBJ/*.o
