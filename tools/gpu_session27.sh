#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facenet.py -m gpu -q -x --timeout=300 -k "repeated or block35 or config2" > gpurun_out/s27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s27_pytest.log
tail -5 gpurun_out/s27_pytest.log
