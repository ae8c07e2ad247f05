#!/bin/bash
# Lab build: libcpfftlog_lab.so with the ablation variants (-DCPF_LAB) and the drivers linked against it
set -e
cd "$(dirname "$0")"
SRC=../../cosmoprimo_b200/csrc
nvcc -O3 -std=c++17 --threads 4 -gencode arch=compute_100a,code=sm_100a -lineinfo -DCPF_LAB -Xcompiler -fPIC -shared \
    -o libcpfftlog_lab.so $SRC/cpf_fftlog.cu $SRC/cpf_peak.cu $SRC/cpf_spline.cu $SRC/cpf_wallish.cu $SRC/cpf_eh.cu
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o pp_driver_lab pp_driver.cu -L. -lcpfftlog_lab -Xlinker -rpath -Xlinker '$ORIGIN'
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o e2e_lab e2e_lab.cu -L../../cosmoprimo_b200 -lcpfftlog -Xlinker -rpath -Xlinker '$ORIGIN/../../cosmoprimo_b200'
