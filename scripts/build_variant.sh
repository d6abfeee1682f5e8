#!/bin/bash
# usage: scripts/build_variant.sh <tag> <nvcc -D flags...>: builds pyfasst_b200/libpyfasst_b200_<tag>.so with
# csrc/estep.cu (or the file named by SRC=) recompiled with the given flags; select it with
# PYFASST_B200_LIB=pyfasst_b200/libpyfasst_b200_<tag>.so (CTA-shape experiments).
set -e
tag=$1; shift
src=${SRC:-estep}
cd "$(dirname "$0")/../pyfasst_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I ../../include "$@" -Xptxas -v -c $src.cu -o build/${src}_$tag.o 2>&1 | grep -A2 "${KERNEL:-estep_stereo_kernelIfLi4}" | grep "spill\|Used" 
objs=$(ls build/*.o | grep -v "_[a-z0-9]*\.o$\|build/$src.o" ; echo build/${src}_$tag.o)
objs=$(for o in build/abi.o build/estep.o build/estep_multi.o build/gemfac.o build/gemm_tc.o build/glue.o build/nmf.o build/nmf_tc.o build/simm.o build/spatial.o build/stft.o build/tc_selftest.o build/viterbi.o build/wf0.o; do if [ "$o" = "build/$src.o" ]; then echo build/${src}_$tag.o; else echo $o; fi; done)
nvcc -shared -o ../libpyfasst_b200_$tag.so $objs -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
ls -la ../libpyfasst_b200_$tag.so
