#!/bin/bash
for st in 0 16384 65536; do
  echo "BN254_STACK=$st"
  BN254_STACK=$st bash tools/gpu_g2check.sh
done
