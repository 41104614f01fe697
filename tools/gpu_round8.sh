#!/bin/bash
TAG=${1:-r01p}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_fpi.py -m gpu -q -x > $O/pytest_fpi.log 2>&1; echo "rc=$?" >> $O/pytest_fpi.log
tail -30 $O/pytest_fpi.log
