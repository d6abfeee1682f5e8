#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wf0_gpu.py -m gpu -q -x -s --timeout=600 2>&1 | tail -15
