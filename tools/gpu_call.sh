#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
bash tools/run_variants.sh
