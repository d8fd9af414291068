#!/bin/bash
mkdir -p gpurun_out
export N_SITES=100000
for pd in 0 1; do
  PDEL=$pd timeout 200 python tools/bench_c4.py 2>&1 | tail -1
  PDEL=$pd NGSDIST_B200_LIB=$PWD/ngsdist_b200/libngsdist_b200_alt.so timeout 200 python tools/bench_c4.py 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_int_path.py tests/test_gpu_block_cache.py tests/test_gpu_shards.py tests/test_gpu_knife.py -q 2>&1 | tail -3
timeout 300 python -m pytest tests/test_cli.py -q -m gpu 2>&1 | tail -3
