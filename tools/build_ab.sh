#!/bin/bash
# Build A/B variants of libapc.so with extra -D flags:  tools/build_ab.sh name "-DFOO=1" ...
# Result: approx_counter_b200/csrc/ab/libapc_<name>.so (git-ignored, travels with gpurun);
# select it with APC_LIB_PATH=approx_counter_b200/csrc/ab/libapc_<name>.so.
set -e
name=$1; shift
cd "$(dirname "$0")/../approx_counter_b200/csrc"
mkdir -p ab/$name/host
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in apc_api apc_comm scan_kernel bitslice_kernel sample_kernels exact_kernels peak_kernels; do
  nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -I. -Ihost "$@" -c $f.cu -o ab/$name/$f.o &
done
for part in 0 1 2 3; do
  nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -I. -Ihost "$@" -DAPC_BS_PART=$part -c bitslice_part.cu -o ab/$name/bspart$part.o &
done
wait
nvcc $ARCH -shared -o ab/libapc_$name.so ab/$name/*.o host/host_util.o host/cli.o host/host_abi.o -cudart static -lgomp -lpthread -ldl
echo built ab/libapc_$name.so
