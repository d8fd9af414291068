#!/bin/bash
# tcgen05 contraction of the called-genotype path: parity tests with the default build, then A/B timing at 5 000 x 100 000 of
# the variants built by tools/build_umma_variants.sh (groups_rawstages_expandedstages)
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_int_path.py tests/test_gpu_block_cache.py tests/test_gpu_configs.py tests/test_gpu_shards.py -q -m gpu 2>&1 | tail -8
for lib in ngsdist_b200/csrc/build/variants/lib_*.so; do
  for pd in 0 1; do
    echo "== $(basename $lib) pairwise_del $pd"
    NGSDIST_B200_LIB=$PWD/$lib N_SITES=100000 PDEL=$pd timeout 60 python tools/bench_c4.py 2>&1 | tail -2
  done
done
} > gpurun_out/umma_planes.log 2>&1
tail -60 gpurun_out/umma_planes.log
