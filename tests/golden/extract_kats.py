#!/usr/bin/env python3
"""Extract the known-answer vectors the reference crate carries in its own unit tests and
write them to tests/golden/reference_kats.json.

Runs ONLY where /root/reference is mounted (the build container); the JSON it writes is
committed so the GPU box and CI never need the reference tree.

Sources (file:line in /root/reference/src/inner_types):
  g1.rs:980-1049   G1 encode (NU) vectors: msg, u, P.x, P.y
  g1.rs:1051-1141  G1 hash   (RO) vectors: msg, u0, u1, Q0, Q1, P
  g2.rs:1038-1225  G2 encode (NU) vectors: canonical little-endian u64 limbs
  g2.rs:1227-1313  G2 hash   (RO) vectors: BE hex
  pairings.rs:387-479  Gt::generator() as Montgomery limbs (decoded here: x * 2^-256 mod p)
  fp2.rs:305-348   G2 generator / twist B as Montgomery limbs
Only DATA (test vectors / constants) is extracted, no code.
"""
import json, re, sys, pathlib

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/inner_types")
OUT = pathlib.Path(__file__).with_name("reference_kats.json")
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
RINV = pow(1 << 256, -1, P)


def unescape(b: str) -> str:
    return bytes(b, "utf-8").decode("unicode_escape")


def section(text, start_pat, end_pat):
    s = re.search(start_pat, text)
    e = re.search(end_pat, text[s.end():])
    return text[s.end(): s.end() + e.start()]


def hex_fields(block):
    return {k: v for k, v in re.findall(r'(\w+)\s*:\s*"([0-9a-fA-F]{64})"', block)}


def split_vectors(body):
    # split on 'TestVector {' occurrences after the struct definition
    parts = body.split("TestVector {")
    return [p for p in parts[1:] if re.search(r'msg:\s*b"', p)]


def msg_of(block):
    return unescape(re.search(r'msg:\s*b"([^"]*)"', block).group(1))


def words_le(ws):
    return sum(int(w, 16) << (64 * i) for i, w in enumerate(ws))


def main():
    g1 = (REF / "g1.rs").read_text()
    g2 = (REF / "g2.rs").read_text()
    pr = (REF / "pairings.rs").read_text()
    f2 = (REF / "fp2.rs").read_text()
    out = {"source": "mikelodder7/bls-bn254 unit-test vectors (gnark-crypto BN254 SVDW suites)"}

    enc = section(g1, r"fn encode\(\) \{", r"for vector in &ENCODE_VECTORS")
    out["g1_encode_dst"] = re.search(r'DST_ENCODE[^=]*=\s*b"([^"]+)"', enc).group(1)
    out["g1_encode"] = []
    for blk in split_vectors(enc):
        f = hex_fields(blk)
        out["g1_encode"].append({"msg": msg_of(blk), "u": f["u"], "p_x": f["p_x"], "p_y": f["p_y"]})

    hs = section(g1, r"fn hash\(\) \{", r"for vector in &HASH_VECTORS")
    out["g1_hash_dst"] = re.search(r'DST_HASH[^=]*=\s*b"([^"]+)"', hs).group(1)
    out["g1_hash"] = []
    for blk in split_vectors(hs):
        f = hex_fields(blk)
        out["g1_hash"].append({"msg": msg_of(blk), **{k: f[k] for k in
                               ("u0", "u1", "q0_x", "q0_y", "q1_x", "q1_y", "p_x", "p_y")}})

    enc2 = section(g2, r"fn encode\(\) \{", r"for vector in (&)?encode_vectors")
    out["g2_encode_dst"] = re.search(r'DST_ENCODE[^=]*=\s*b"([^"]+)"', enc2).group(1)
    out["g2_encode"] = []
    for blk in split_vectors(enc2):
        limbs = re.findall(r"0x([0-9a-fA-F]{1,16})\b", blk)
        assert len(limbs) == 16, len(limbs)
        v = [words_le(limbs[i * 4:(i + 1) * 4]) for i in range(4)]
        out["g2_encode"].append({"msg": msg_of(blk), "x_c0": "%064x" % v[0], "x_c1": "%064x" % v[1],
                                 "y_c0": "%064x" % v[2], "y_c1": "%064x" % v[3]})

    hs2 = section(g2, r"fn hash\(\) \{", r"for vector in (&)?hash_vectors")
    out["g2_hash_dst"] = re.search(r'DST_HASH[^=]*=\s*b"([^"]+)"', hs2).group(1)
    out["g2_hash"] = []
    for blk in split_vectors(hs2):
        hx = re.findall(r'from_be_hex\("([0-9a-fA-F]{64})"\)', blk)
        assert len(hx) == 4
        out["g2_hash"].append({"msg": msg_of(blk), "x_c0": hx[0], "x_c1": hx[1], "y_c0": hx[2], "y_c1": hx[3]})

    gen = section(pr, r"fn generator\(\) -> Self \{\s*// pairing\(G1Affine::generator\(\), G2Affine::generator\(\)\)",
                  r"\n    fn ")
    limbs = re.findall(r"0x([0-9a-fA-F]{16})", gen)
    assert len(limbs) == 48, len(limbs)
    out["gt_generator"] = ["%064x" % (words_le(limbs[i * 4:(i + 1) * 4]) * RINV % P) for i in range(12)]
    out["gt_generator_order"] = "c0.c0.c0 c0.c0.c1 c0.c1.c0 c0.c1.c1 c0.c2.c0 c0.c2.c1 c1.c0.c0 ... (Gt::to_repr order)"

    consts = section(f2, r"pub const GEN_X: Self", r"pub const B3")
    limbs = re.findall(r"0x([0-9a-fA-F]{16})", consts)
    assert len(limbs) == 24
    vals = ["%064x" % (words_le(limbs[i * 4:(i + 1) * 4]) * RINV % P) for i in range(6)]
    out["g2_generator"] = {"x_c0": vals[0], "x_c1": vals[1], "y_c0": vals[2], "y_c1": vals[3]}
    out["g2_b"] = {"c0": vals[4], "c1": vals[5]}

    for k in ("g1_encode", "g1_hash", "g2_encode", "g2_hash"):
        assert len(out[k]) == 5, (k, len(out[k]))
    OUT.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", OUT, {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
