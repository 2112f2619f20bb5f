#!/bin/bash
# A/B builds of the library for kernel experiments:  tools/build_variant.sh NAME FILE.cu=ALT.cu [-DFLAG ...]
# compiles ALT.cu in place of aruco3_b200/csrc/FILE.cu (or the same file with extra -D flags), links it with the other objects
# of the regular build into variants/libNAME.so (git-ignored, travels with gpurun); load it with A3_LIB_PATH=variants/libNAME.so
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; SPEC=$2; shift 2
FILE=${SPEC%%=*}; ALT=${SPEC#*=}
CS=$ROOT/aruco3_b200/csrc
make -s -C $CS >/dev/null
mkdir -p $ROOT/variants/obj
FM=""
case $FILE in k2_decode.cu|k3_contours.cu|k4_pose.cu) FM="-fmad=false";; esac
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-ffp-contract=off -Xptxas -v $FM \
    -I$CS -I$ROOT/include "$@" -c $ALT -o $ROOT/variants/obj/$NAME.o 2> $ROOT/variants/obj/$NAME.ptxas.log || { cat $ROOT/variants/obj/$NAME.ptxas.log; exit 1; }
OBJS=""
for o in $CS/build/*.o; do
  [ "$(basename $o .o)" = "$(basename $FILE .cu)" ] && continue
  OBJS="$OBJS $o"
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/lib$NAME.so $ROOT/variants/obj/$NAME.o $OBJS -lpthread
echo "built variants/lib$NAME.so"
