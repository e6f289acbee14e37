#!/usr/bin/env python3
"""Freeze the EXECUTED multiplier work per item of every hot kernel body (tests/hostsim build of the device headers,
fp.cuh: BN_COUNT_OPS) next to the canonical counts of SURVEY.md 8(d).  Writes profiles/executed_ops.json; bench.py
reads it for `frac_executed`, tests/test_bench_contract.py checks it is in sync with the sources.

  python tools/count_ops.py [--check]
"""
import ctypes as C
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "profiles" / "executed_ops.json"
M = 264  # INT32 multiply-add operations of one Montgomery product
CANONICAL_M = {"k_hash_g1": 2204, "k_hash_g2": 6400, "k_miller1": 8610, "k_miller2_fixed": 12384, "k_miller2": 14860, "k_miller_twin": 14860,
               "k_final_exp": 8427, "k_g2_require_subgroup": None}
KINDS = {"k_hash_g1": 0, "k_hash_g2": 1, "k_miller1": 2, "k_miller2_fixed": 3, "k_miller2": 4, "k_final_exp": 5, "k_g2_require_subgroup": 6}


def measure():
    from oracle import orc
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "tests" / "hostsim")])
    sim = C.CDLL(str(ROOT / "tests" / "hostsim" / "libhostsim.so"))
    sim.sim_count_item.restype = C.c_ulonglong
    sim.sim_count_item.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    dst = b"BLS_SIG_BN254G1_XMD:SHA-256_SVDW_RO_NUL_"
    msgs = [bytes([i]) * 32 for i in range(1, 9)]
    g1 = orc.hash_to_g1(msgs, dst)
    g2 = orc.hash_to_g2(msgs, dst)
    out = {}
    for name, kind in KINDS.items():
        vals = []
        for i in range(4):
            vals.append(int(sim.sim_count_item(kind, g1[64 * i:64 * i + 64], g2[128 * i:128 * i + 128], g1[64 * (i + 4):64 * (i + 5)],
                                               g2[128 * (i + 4):128 * (i + 5)], msgs[i], 32, dst, len(dst))))
        ops = sum(vals) / len(vals)
        out[name] = {"imad_ops_per_item": ops, "M_equivalent": round(ops / M, 1), "canonical_M": CANONICAL_M[name], "spread": max(vals) - min(vals)}
    out["k_miller_twin"] = dict(out["k_miller2"], canonical_M=CANONICAL_M["k_miller_twin"])
    return out


def main():
    got = measure()
    if "--check" in sys.argv:
        have = json.loads(OUT.read_text())["kernels"]
        bad = [k for k in got if abs(have.get(k, {}).get("imad_ops_per_item", -1) - got[k]["imad_ops_per_item"]) > 0.002 * got[k]["imad_ops_per_item"]]
        if bad:
            raise SystemExit("profiles/executed_ops.json is stale for: %s (run tools/count_ops.py)" % ", ".join(bad))
        print("executed_ops.json in sync")
        return
    doc = {"unit": "INT32 multiply-add operations per item (one IMAD.WIDE = 2), multiplier routines only, wire codecs excluded",
           "M": M, "how": "tests/hostsim build of bls_bn254_b200/csrc/*.cuh, counter in fp.cuh (BN_COUNT_OPS), mean of 4 items",
           "kernels": got}
    OUT.write_text(json.dumps(doc, indent=1) + "\n")
    for k, v in got.items():
        print("%-24s executed %8.1f M   canonical %s" % (k, v["M_equivalent"], v["canonical_M"]))


if __name__ == "__main__":
    main()
