#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q --durations=5 2>&1 | tail -25
