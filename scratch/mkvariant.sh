#!/bin/bash
# scratch/mkvariant.sh NAME [-DFLAG=..]...  ->  scratch/variants/NAME.so  (kernel experiments; load with B200MEL_LIB=...)
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o scratch/variants/$name.so audio_transformers_b200/csrc/b200mel.cu
