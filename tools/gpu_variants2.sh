#!/bin/bash
# bench (and a parity subset) for tuning builds under build/: BN254_SO selects the library
mkdir -p gpurun_out
for so in "$@"; do
  name=$(basename $so .so)
  export BN254_SO=$PWD/$so; export BN254_LANES=${FORCE_LANES:-0}
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_verify_vs_oracle or test_pairing_vs_oracle or test_gt_generator" > gpurun_out/pytest_$name.log 2>&1; echo "$name pytest rc=$? $(tail -1 gpurun_out/pytest_$name.log)"
  timeout 600 python bench.py --steps 5 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name bench rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('$name value %.0f'%d['value'], ' '.join('%s=%.2f'%(k.split('[')[0],v['ms_per_launch']) for k,v in d['roofline']['kernels'].items()))
"
done
