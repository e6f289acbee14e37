#!/bin/bash
# first contact with the B200: smoke, IMAD peak, the GPU parity suite
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; which cargo rustc >> gpurun_out/nproc.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python - > gpurun_out/imad.log 2>&1 <<'PY'
import bls_bn254_b200 as b
c = b.Context([0])
for wide in (False, True):
    for it in (2000, 20000):
        ops, ms = c.imad_peak(wide, it)
        print("wide=%d iters=%d  %.3f T IMAD-ops/s  %.3f ms" % (wide, it, ops / 1e12, ms))
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/smoke.log gpurun_out/imad.log gpurun_out/pytest_gpu.log
