#!/bin/bash
# two GPUs: the NCCL exchange inside the library (multi-device ctx test + torchrun bench with the c4 / c5 sub-records)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device or nccl" > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_n2.log
NG=${NG:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 5 --warmup 3 "$@" > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n$NG.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n$NG.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],'n_gpus',d['n_gpus'])
for k in ('c3','pairing','c4','c5'): print(k, json.dumps(d.get(k)))
"
