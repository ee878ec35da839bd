#!/bin/bash
o=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
bash tools/run_variants.sh 2>&1
