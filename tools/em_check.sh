#!/bin/bash
python -m pytest tests/test_gpu_em.py tests/test_gpu_configs.py::test_c1_shape_em_default -x -q -m gpu 2>&1 | tail -3
python tools/bench_em.py
