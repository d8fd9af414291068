#!/bin/bash
python - <<'PY'
import oracle
with open("/tmp/in.bin","wb") as fh:
    for s0 in range(0,100000,10000): oracle.synth_raw(20251018,0.0,500,10000,site0=s0).tofile(fh)
PY
for k in 1 2; do NGSD_CLI_TIMING=1 ngsdist_b200/bin/ngsDist --geno /tmp/in.bin --n_ind 500 --n_sites 100000 --out /tmp/o.dist --n_threads 16 --verbose 0 --probs --indep_geno --evol_model 2 2>&1 | grep timing; echo; done
python -m pytest "tests/test_gpu_edges.py" -x -q -m gpu 2>&1 | tail -30
