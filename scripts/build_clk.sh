#!/bin/bash
# development: second library with -DGLORIA_PHASE_CLOCKS (load it with GLORIA_B200_LIB=...)
cd "$(dirname "$0")/../gloria_nlp_project_b200" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DGLORIA_PHASE_CLOCKS -o libgloria_b200_clk.so csrc/*.cu -lcublas
