// Unity translation unit of libola_gpu.so: one device-code module so that __constant__ tables and
// __forceinline__ field arithmetic are visible to every kernel without relocatable device code.
// Built by olavm_b200/build.py:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC
#include "ntt.cu"
#include "poseidon.cu"
#include "blake3.cu"
#include "batch.cu"
#include "fri.cu"
#include "stark.cu"
#include "generation.cu"
#include "lookup.cu"
#include "gen_tables.cu"
#include "nccl_comm.cu"
#include "api.cu"
