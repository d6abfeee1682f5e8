#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sourcefilter_gpu.py -m gpu -q -x --timeout=600 2>&1 | tail -25
