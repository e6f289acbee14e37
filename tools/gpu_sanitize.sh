#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/g2small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import bls_bn254_b200 as b
from oracle import orc
c = b.Context([0])
msgs = [b"abc", b"", b"xyz123"]
dst = b"QUUX-V01-CS02-with-BN254G2_XMD:SHA-256_SVDW_NU_"
out = c.encode_to_g2(msgs, dst)
print("match oracle:", out == orc.hash_to_g2(msgs, dst, 1))
PY
BN254_SO=$PWD/build/variants/libbase.so timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/g2small.py > gpurun_out/sanitize.log 2>&1
echo rc=$?; tail -40 gpurun_out/sanitize.log
