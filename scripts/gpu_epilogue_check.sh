#!/bin/bash
# short validation of the lean K2 epilogue: single-CTA batched parity tests, then kernel times on the c2 shape
mkdir -p gpurun_out
timeout 118 python -m pytest tests/test_gpu_batched.py -q -x -k "single_cta or single_pass or batch_mode_never" > gpurun_out/pytest_epilogue.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_epilogue.log
QUICK=1 timeout 40 python scripts/dbg_passes.py > gpurun_out/passes_epilogue.log 2>&1
echo "timing rc=$?"; cat gpurun_out/passes_epilogue.log
