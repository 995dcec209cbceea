#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_tests3.log
