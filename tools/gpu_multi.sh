#!/bin/bash
# N-GPU runs of bench.py exactly as the driver launches them
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
for w in verify verify_rand aggregate; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $w > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err
  echo "$w N=$N rc=$?"; tail -c 400 gpurun_out/bench_${w}_n$N.err
  python - $w $N <<'PY'
import json,sys
w,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open('gpurun_out/bench_%s_n%s.json'%(w,n)).read().strip().splitlines()[-1])
    print(w, "n_gpus", d['n_gpus'], "value=%.0f ms=%.2f e2e=%.0f"%(d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e: print("parse fail", e)
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 | tail -c 300
timeout 600 python -m pytest tests -m gpu -q -k "multi_device" 2>&1 | tail -3
# C5: 2^24 tuples through the host-buffer call of ONE process driving all N GPUs (one host thread per device)
C5_GPUS=$N bash tools/gpu_c5.sh
