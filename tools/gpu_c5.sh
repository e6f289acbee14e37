#!/bin/bash
# C5 (2^24 tuples through the host API) + refreshed bench lines for the randomised workload
mkdir -p gpurun_out
timeout 1200 python tools/bench_c5.py --gpus ${C5_GPUS:-1} --log2n ${C5_LOG2N:-24} > gpurun_out/c5_n${C5_GPUS:-1}.json 2> gpurun_out/c5_n${C5_GPUS:-1}.err
echo "c5 rc=$?"; cat gpurun_out/c5_n${C5_GPUS:-1}.json | cut -c1-400; tail -3 gpurun_out/c5_n${C5_GPUS:-1}.err
