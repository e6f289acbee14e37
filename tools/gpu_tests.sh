#!/bin/bash
# GPU check of the parity suite: smoke, then pytest -m gpu (log + per-test durations under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=15 "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
