#!/bin/bash
# builds libngsdist_b200 with several (expander groups, raw stages, expanded stages) settings of dist_umma.cu for one A/B GPU run
set -e
cd "$(dirname "$0")/../ngsdist_b200/csrc"
make -s >/dev/null
mkdir -p build/variants
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-fvisibility=hidden --compiler-bindir /usr/bin/g++"
for v in "$@"; do
  IFS=_ read g r e <<< "$v"
  /usr/local/cuda/bin/nvcc $NVFLAGS -DNGSD_UMMA_GROUPS=$g -DNGSD_UMMA_KRAW=$r -DNGSD_UMMA_KEXP=$e -c dist_umma.cu -o build/variants/dist_umma_$v.o
  objs=$(ls build/*.o | grep -v dist_umma.o)
  /usr/local/cuda/bin/nvcc -shared -o build/variants/lib_$v.so $objs build/variants/dist_umma_$v.o -cudart static -lnccl
  echo built $v
done
