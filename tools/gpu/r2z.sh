#!/bin/sh
# round 2, final state: the whole -m gpu suite (the r2y call stopped at the fuzz harness's coded-row tolerance)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2z_pytest.txt 2>&1
tail -4 gpurun_out/r2z_pytest.txt
