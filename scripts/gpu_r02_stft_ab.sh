#!/bin/bash
# CTA shape of the STFT kernel: threads per CTA (build variants) x frames per batch
mkdir -p gpurun_out
run() { PYFASST_B200_LIB=$1 PYFASST_STFT_NB=$2 timeout 100 python scripts/micro/stft_time.py 2>&1 | tail -1; }
{
timeout 100 python scripts/micro/stft_time.py 2>&1 | tail -1
run pyfasst_b200/libpyfasst_b200_t512.so 2
run pyfasst_b200/libpyfasst_b200_t512.so 4
run pyfasst_b200/libpyfasst_b200_t1024.so 4
run pyfasst_b200/libpyfasst_b200_t1024.so 8
} | tee gpurun_out/stft_shape_ab.txt
