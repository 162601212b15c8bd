#!/bin/bash
# round 2, second session, run 11: the full-size property test of a bf16 row store (10M x 768).
mkdir -p gpurun_out/r2b11
timeout 120 python -m pytest tests/test_gpu_fullsize.py -q -x -k bf16 > gpurun_out/r2b11/pytest.log 2>&1; echo "fullsize bf16 rc=$?"; tail -3 gpurun_out/r2b11/pytest.log
