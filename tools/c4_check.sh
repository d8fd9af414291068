#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
N_SITES=100000 python tools/bench_c4.py
N_SITES=100000 PDEL=0 python tools/bench_c4.py
N_IND=2000 N_SITES=50000 FORCE_FP64=1 python tools/bench_c4.py
