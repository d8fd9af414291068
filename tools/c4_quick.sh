#!/bin/bash
python -m pytest tests/test_gpu_int_path.py -x -q -m gpu 2>&1 | tail -2
N_SITES=100000 PDEL=0 python tools/bench_c4.py
N_SITES=100000 PDEL=1 python tools/bench_c4.py
