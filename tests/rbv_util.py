"""Helpers shared by the CPU and GPU tests of the randomised batch verification."""
import hashlib

from oracle import pymodel as pm


def _rbv_scalar(seed, i, bits=64):
    """r_i = a + b lambda mod r (include/bls_bn254_b200.h), as an integer"""
    return pm.rbv_scalar(seed, i, bits)


def _neg_g1(pt):
    return pt[:32] + ((pm.P - int.from_bytes(pt[32:], "big")) % pm.P).to_bytes(32, "big")


def rbv_forgery(orc, sigs, seed, bits=64, i=0, j=1):
    """sig_i += r_j D, sig_j -= r_i D (D = the G1 generator): both tuples become invalid, yet sum_k r_k sig_k is
    unchanged, so a verifier whose scalars were predicted accepts.  Pins the scalar derivation bit-exactly."""
    ri, rj = _rbv_scalar(seed, i, bits), _rbv_scalar(seed, j, bits)
    d = orc.g1_generator()
    di = orc.g1_mul(d, rj.to_bytes(32, "big"))
    dj = _neg_g1(orc.g1_mul(d, ri.to_bytes(32, "big")))
    out = bytearray(sigs)
    out[64 * i:64 * (i + 1)] = orc.g1_sum(sigs[64 * i:64 * (i + 1)] + di)
    out[64 * j:64 * (j + 1)] = orc.g1_sum(sigs[64 * j:64 * (j + 1)] + dj)
    return bytes(out)
