#!/bin/bash
# correctness of every build/variants/lib*.so on the min-sig / min-pk verify tests, then timing
for so in build/variants/lib*.so; do
  v=$(basename $so .so)
  BN254_SO=$PWD/$so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "verify_vs_oracle or ragged or pairing_vs_oracle" 2>&1 | tail -1 | sed "s/^/$v: /"
done
