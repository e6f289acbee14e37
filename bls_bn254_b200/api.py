"""Host-side mirror of the reference crate's `inner_types` API, in batch form, over the C ABI.

Reference function  ->  here (all byte formats are the reference's own, see include/bls_bn254_b200.h)
  G1Projective::hash / ::encode (g1.rs:910-928)            Context.hash_to_g1 / encode_to_g1
  G2Projective::hash / ::encode (g2.rs:919-936)            Context.hash_to_g2 / encode_to_g2
  pairing (pairings.rs:760-802) + Gt::to_repr              Context.pairing
  multi_miller_loop (pairings.rs:808-857)                  Context.multi_miller_loop
  MillerLoopResult::final_exponentiation (50-178)          Context.final_exponentiation
  MillerLoopResult + MillerLoopResult (37-43)              Context.fp12_product
  to_compressed / from_compressed (g1.rs:283-328, g2.rs)   Context.g1_compress / g1_decompress / g2_...
  G1Projective * Scalar (g1.rs:821-841, g2.rs:866-886)     Context.g1_mul / g2_mul
  verify / aggregate_verify (absent upstream) Context.verify_batch / aggregate_verify
Errors: malformed call-level input raises Bn254Error (error.rs:5-10); per-item problems come back
as status bytes, never as exceptions.
"""
import ctypes as C

from . import _native

MIN_SIG, MIN_PK = 0, 1
CHECK_ON_CURVE = 1
CHECK_SUBGROUP = 2


class Bn254Error(Exception):
    """error.rs:5-10 (InvalidScalarBytes / InvalidG1Bytes / InvalidG2Bytes / InvalidGtBytes) + CUDA errors."""
    NAMES = {-1: "InvalidScalarBytes", -2: "InvalidG1Bytes", -3: "InvalidG2Bytes", -4: "InvalidGtBytes", -10: "BadArgument", -20: "NoCudaDevice"}

    def __init__(self, code, msg=""):
        self.code = code
        kind = self.NAMES.get(code, "NcclError" if code <= -2000 else "CudaError")
        super().__init__("%s (%d)%s" % (kind, code, (": " + msg) if msg else ""))


def pack_msgs(msgs):
    """list[bytes] -> (blob, (n+1) x u64 offsets)"""
    off = (C.c_uint64 * (len(msgs) + 1))()
    pos = 0
    for i, m in enumerate(msgs):
        off[i] = pos
        pos += len(m)
    off[len(msgs)] = pos
    return b"".join(msgs), off


def _msgs(msgs):
    """list[bytes] or an already packed (blob, offsets) pair -> (blob, offsets, n); a packed pair is checked against its
    blob (the C ABI has no blob length: offsets[n] is what bounds the reads)."""
    if not isinstance(msgs, tuple):
        blob, off = pack_msgs(msgs)
        return blob, off, len(off) - 1
    blob, off = msgs
    n = len(off) - 1
    if n < 0:
        raise Bn254Error(-10, "offsets must hold n + 1 entries")
    if isinstance(blob, (bytes, bytearray)) and n >= 0 and int(off[n]) > len(blob):
        raise Bn254Error(-10, "offsets[n] exceeds the message blob")
    return blob, off, n


def _ptr(x):
    """bytes / bytearray / ctypes buffer / numpy array / int address / None -> void*"""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(x)) if isinstance(x, bytearray) else C.c_char_p(x), C.c_void_p)
    if hasattr(x, "ctypes"):  # numpy
        return C.c_void_p(x.ctypes.data)
    return C.cast(x, C.c_void_p)


class PreparedKeys:
    """G2Prepared (pairings.rs:609-660) for a set of min-sig public keys: the 88 line triples of every key, resident on the
    GPU(s) of the Context that made it (Context.g2_prepare).  Free with close() (or let it go out of scope)."""

    def __init__(self, ctx, handle, n):
        self._ctx, self._h, self.n = ctx, handle, n

    def close(self):
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            self._ctx._L.bn254_g2_prepared_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.n

    def export(self, first=0, count=None):
        """(tables, status): 88 x 192 bytes per key (88 x 3 Fp2 as c0 || c1 big-endian), one status byte per key"""
        count = self.n - first if count is None else count
        out, st = C.create_string_buffer(max(1, 88 * 192 * count)), C.create_string_buffer(max(1, count))
        self._ctx._ck(self._ctx._L.bn254_g2_prepared_export(self._ctx._h, self._h, first, count, _ptr(out), _ptr(st)))
        return out.raw[:88 * 192 * count], st.raw[:count]


class Context:
    """bn254_ctx: owns streams and device scratch on one or more GPUs of this process."""

    def __init__(self, devices=None):
        self._L = _native.lib()
        if devices is None:
            devices = [0]
        if isinstance(devices, int):
            devices = [devices]
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._L.bn254_ctx_create(C.byref(h), arr, len(devices))
        if rc:
            raise Bn254Error(rc, "bn254_ctx_create failed (no CUDA device? this library has no CPU fallback)")
        self._h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None):
            self._L.bn254_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise Bn254Error(rc, (self._L.bn254_last_error(self._h) or b"").decode())

    @property
    def launch_count(self):
        return int(self._L.bn254_launch_count(self._h))

    @property
    def stream_ptr(self):
        return int(self._L.bn254_stream(self._h) or 0)

    def sync(self):
        self._ck(self._L.bn254_sync(self._h))

    def set_option(self, name, value):
        """"lanes" (0 per launch / 1 one thread per pairing / 2 lane pair / 3 slot machine / 4 wide slot machine), "wide_max", "chunk", "split", "sk_msm_min":
        include/bls_bn254_b200.h"""
        self._ck(self._L.bn254_ctx_set_option(self._h, name.encode(), int(value)))

    # ---- NCCL communicator of a one-process-per-GPU job (the aggregate path's single exchange) -----------------
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        rc = self._L.bn254_comm_unique_id(_ptr(buf))
        if rc:
            raise Bn254Error(rc, "bn254_comm_unique_id (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init_rank(self, unique_id, nranks, rank):
        if len(unique_id) != 128:
            raise Bn254Error(-10, "unique id must be 128 bytes")
        self._ck(self._L.bn254_comm_init_rank(self._h, _ptr(bytes(unique_id)), int(nranks), int(rank)))

    @property
    def comm_nranks(self):
        return int(self._L.bn254_comm_nranks(self._h))

    def profile(self, on=True):
        self._ck(self._L.bn254_profile_enable(self._h, 1 if on else 0))

    def profile_drain(self):
        """[(kernel name, items, milliseconds)] for every launch since the last drain (all devices, device order)."""
        buf = C.create_string_buffer(1 << 20)
        self._ck(self._L.bn254_profile_drain(self._h, _ptr(buf), len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            name, items, ms = line.split()
            out.append((name, int(items), float(ms)))
        return out

    # ---- hash to curve ------------------------------------------------------------------------
    def _hash(self, fn, rec, msgs, dst, want_status):
        blob, off, n = _msgs(msgs)
        out, st = C.create_string_buffer(rec * n), C.create_string_buffer(n)
        self._ck(fn(self._h, _ptr(blob), _ptr(off), n, _ptr(dst), len(dst), _ptr(out), _ptr(st)))
        return (out.raw, st.raw) if want_status else out.raw

    def hash_to_g1(self, msgs, dst, want_status=False):
        return self._hash(self._L.bn254_hash_to_g1, 64, msgs, dst, want_status)

    def encode_to_g1(self, msgs, dst, want_status=False):
        return self._hash(self._L.bn254_encode_to_g1, 64, msgs, dst, want_status)

    def hash_to_g2(self, msgs, dst, want_status=False):
        return self._hash(self._L.bn254_hash_to_g2, 128, msgs, dst, want_status)

    def encode_to_g2(self, msgs, dst, want_status=False):
        return self._hash(self._L.bn254_encode_to_g2, 128, msgs, dst, want_status)

    # ---- pairing engine -----------------------------------------------------------------------
    def pairing(self, g1, g2, want_status=False):
        n = len(g1) // 64
        if len(g1) != 64 * n or len(g2) != 128 * n:
            raise Bn254Error(-2 if len(g1) != 64 * n else -3, "length mismatch")
        out, st = C.create_string_buffer(384 * n), C.create_string_buffer(n)
        self._ck(self._L.bn254_pairing_batch(self._h, _ptr(g1), _ptr(g2), n, _ptr(out), _ptr(st)))
        return (out.raw, st.raw) if want_status else out.raw

    def miller_loop(self, g1, g2, want_status=False):
        n = len(g1) // 64
        if len(g1) != 64 * n or len(g2) != 128 * n:
            raise Bn254Error(-2 if len(g1) != 64 * n else -3, "length mismatch")
        out, st = C.create_string_buffer(384 * n), C.create_string_buffer(n)
        self._ck(self._L.bn254_miller_loop_batch(self._h, _ptr(g1), _ptr(g2), n, _ptr(out), _ptr(st)))
        return (out.raw, st.raw) if want_status else out.raw

    def multi_miller_loop(self, g1, g2):
        n = len(g1) // 64
        if len(g1) != 64 * n or len(g2) != 128 * n:
            raise Bn254Error(-2 if len(g1) != 64 * n else -3, "length mismatch")
        out = C.create_string_buffer(384)
        self._ck(self._L.bn254_multi_miller_loop(self._h, _ptr(g1), _ptr(g2), n, _ptr(out)))
        return out.raw

    def final_exponentiation(self, ml, want_status=False):
        n = len(ml) // 384
        if len(ml) != 384 * n:
            raise Bn254Error(-4, "length mismatch")
        out, st = C.create_string_buffer(384 * n), C.create_string_buffer(n)
        self._ck(self._L.bn254_final_exponentiation(self._h, _ptr(ml), n, _ptr(out), _ptr(st)))
        return (out.raw, st.raw) if want_status else out.raw

    def fp12_product(self, elems):
        n = len(elems) // 384
        if len(elems) != 384 * n:
            raise Bn254Error(-4, "length mismatch")
        out = C.create_string_buffer(384)
        self._ck(self._L.bn254_fp12_product(self._h, _ptr(elems), n, _ptr(out)))
        return out.raw

    # ---- BLS predicates -----------------------------------------------------------------------
    @staticmethod
    def _recs(variant):
        return (128, 64) if variant == MIN_SIG else (64, 128)

    def verify_batch(self, variant, pks, msgs, sigs, dst, flags=0):
        blob, off, n = _msgs(msgs)
        rpk, rsig = self._recs(variant)
        if len(pks) != rpk * n:
            raise Bn254Error(-3 if variant == MIN_SIG else -2, "public-key buffer length")
        if len(sigs) != rsig * n:
            raise Bn254Error(-2 if variant == MIN_SIG else -3, "signature buffer length")
        ok = C.create_string_buffer(n)
        self._ck(self._L.bn254_verify_batch(self._h, variant, _ptr(pks), _ptr(blob), _ptr(off), _ptr(sigs), n, _ptr(dst), len(dst), flags, _ptr(ok)))
        return ok.raw

    def fast_aggregate_verify_batch(self, variant, pks, key_counts, msgs, agg_sigs, dst, flags=0):
        """FastAggregateVerify for many groups: group g = the next key_counts[g] public keys of `pks`, message g, aggregate
        signature g.  One accept byte per group."""
        blob, off, n = _msgs(msgs)
        rpk, rsig = self._recs(variant)
        if len(key_counts) != n or len(agg_sigs) != rsig * n or len(pks) != rpk * sum(key_counts):
            raise Bn254Error(-10, "one key count and one aggregate signature per message; pks must hold sum(key_counts) keys")
        koff = (C.c_uint64 * (n + 1))()
        for i, k in enumerate(key_counts):
            koff[i + 1] = koff[i] + k
        ok = C.create_string_buffer(max(1, n))
        self._ck(self._L.bn254_fast_aggregate_verify_batch(self._h, variant, _ptr(pks), _ptr(koff), _ptr(blob), _ptr(off), _ptr(agg_sigs), n, _ptr(dst),
                                                           len(dst), flags, _ptr(ok)))
        return ok.raw[:n]

    def verify_same_key_randomized(self, variant, pks, msg_counts, msgs, sigs, dst, seed, rand_bits=64, flags=0):
        """Same-key randomised verification: key k owns the next msg_counts[k] (message, signature) tuples and gets ONE accept
        byte (one Miller loop and one final exponentiation per KEY; per tuple a hash and two short scalar multiplications).
        `seed` = 32 bytes the signers cannot predict."""
        blob, off, n = _msgs(msgs)
        rpk, rsig = self._recs(variant)
        nk = len(msg_counts)
        if len(pks) != rpk * nk:
            raise Bn254Error(-3 if variant == MIN_SIG else -2, "public-key buffer length")
        if sum(msg_counts) != n or len(sigs) != rsig * n:
            raise Bn254Error(-10, "msg_counts must add up to the number of messages, one signature per message")
        if len(seed) != 32:
            raise Bn254Error(-10, "seed must be 32 bytes")
        koff = (C.c_uint64 * (nk + 1))()
        for i, k in enumerate(msg_counts):
            koff[i + 1] = koff[i] + k
        ok = C.create_string_buffer(max(1, nk))
        self._ck(self._L.bn254_verify_same_key_randomized(self._h, variant, _ptr(pks), _ptr(koff), nk, _ptr(blob), _ptr(off), _ptr(sigs), _ptr(dst),
                                                          len(dst), flags, _ptr(seed), rand_bits, _ptr(ok)))
        return ok.raw[:nk]

    def verify_same_key_randomized_dev(self, variant, pks, key_offsets_host, n_keys, msgs, offsets, sigs, dst, dst_len, flags, seed, rand_bits, ok):
        self._ck(self._L.bn254_verify_same_key_randomized_dev(self._h, variant, pks, key_offsets_host, n_keys, msgs, offsets, sigs, dst, dst_len, flags,
                                                              seed, rand_bits, ok))

    def g2_prepare(self, pks, flags=0):
        """G2Prepared::from for n min-sig public keys (n x 128 bytes): prepare once, verify many (verify_batch_prepared).
        `flags` = validation of the keys (CHECK_ON_CURVE / CHECK_SUBGROUP); identity keys are unusable either way."""
        n = len(pks) // 128
        if len(pks) != 128 * n:
            raise Bn254Error(-3, "public-key buffer length")
        h = C.c_void_p()
        self._ck(self._L.bn254_g2_prepare(self._h, _ptr(pks), n, flags, C.byref(h)))
        return PreparedKeys(self, h, n)

    def verify_batch_prepared(self, prepared, key_index, msgs, sigs, dst, flags=0):
        """n accept bytes: tuple i is (prepared key key_index[i], msg i, sig i), min-sig.  Same verdicts as verify_batch."""
        blob, off, n = _msgs(msgs)
        if len(key_index) != n:
            raise Bn254Error(-10, "one key index per tuple")
        if len(sigs) != 64 * n:
            raise Bn254Error(-2, "signature buffer length")
        ki = key_index if hasattr(key_index, "ctypes") else (C.c_uint32 * n)(*key_index)
        ok = C.create_string_buffer(max(1, n))
        self._ck(self._L.bn254_verify_batch_prepared(self._h, prepared._h, _ptr(ki), _ptr(blob), _ptr(off), _ptr(sigs), n, _ptr(dst), len(dst), flags, _ptr(ok)))
        return ok.raw[:n]

    def verify_batch_prepared_dev(self, prepared, key_index, msgs, offsets, sigs, n, dst, dst_len, flags, ok):
        self._ck(self._L.bn254_verify_batch_prepared_dev(self._h, prepared._h, key_index, msgs, offsets, sigs, n, dst, dst_len, flags, ok))

    def verify_batch_randomized(self, variant, pks, msgs, sigs, dst, seed, group=256, rand_bits=64, flags=0):
        """One accept byte per consecutive group of `group` tuples (small-exponent batch test).
        `seed` = 32 bytes the signers cannot predict (fresh CSPRNG output per call)."""
        blob, off, n = _msgs(msgs)
        rpk, rsig = self._recs(variant)
        if len(pks) != rpk * n:
            raise Bn254Error(-3 if variant == MIN_SIG else -2, "public-key buffer length")
        if len(sigs) != rsig * n:
            raise Bn254Error(-2 if variant == MIN_SIG else -3, "signature buffer length")
        if len(seed) != 32:
            raise Bn254Error(-10, "seed must be 32 bytes")
        ng = (n + group - 1) // group if group > 0 else 0
        ok = C.create_string_buffer(max(ng, 1))
        self._ck(self._L.bn254_verify_batch_randomized(self._h, variant, _ptr(pks), _ptr(blob), _ptr(off), _ptr(sigs), n, _ptr(dst), len(dst), flags,
                                                       _ptr(seed), rand_bits, group, _ptr(ok)))
        return ok.raw[:ng]

    def aggregate_verify(self, variant, pks, msgs, agg_sig, dst, flags=0):
        blob, off, n = _msgs(msgs)
        rpk, rsig = self._recs(variant)
        if len(pks) != rpk * n:
            raise Bn254Error(-3 if variant == MIN_SIG else -2, "public-key buffer length")
        if len(agg_sig) != rsig:
            raise Bn254Error(-2 if variant == MIN_SIG else -3, "signature length")
        ok = C.create_string_buffer(1)
        self._ck(self._L.bn254_aggregate_verify(self._h, variant, _ptr(pks), _ptr(blob), _ptr(off), n, _ptr(agg_sig), _ptr(dst), len(dst), flags, _ptr(ok)))
        return ok.raw[0]

    def aggregate_verify_dist(self, variant, pks_shard, msgs_shard, agg_sig, dst, flags=0):
        """Every rank of a one-process-per-GPU job calls this with ITS shard; agg_sig is read on rank 0.  One NCCL
        all-gather of 400 bytes per rank, one post-gather kernel; returns the job's verdict on every rank."""
        blob, off, n = _msgs(msgs_shard)
        rpk, rsig = self._recs(variant)
        if len(pks_shard) != rpk * n:
            raise Bn254Error(-3 if variant == MIN_SIG else -2, "public-key buffer length")
        if agg_sig is not None and len(agg_sig) != rsig:
            raise Bn254Error(-2 if variant == MIN_SIG else -3, "signature length")
        ok = C.create_string_buffer(1)
        self._ck(self._L.bn254_aggregate_verify_dist(self._h, variant, _ptr(pks_shard), _ptr(blob), _ptr(off), n, _ptr(agg_sig), _ptr(dst), len(dst),
                                                     flags, _ptr(ok)))
        return ok.raw[0]

    def aggregate_partial(self, variant, pks, msgs, agg_sig, dst, flags=0):
        """One shard: (384-byte Fp12 partial product, bad flag).  agg_sig only on one shard."""
        blob, off, n = _msgs(msgs)
        rpk, rsig = self._recs(variant)
        if len(pks) != rpk * n:
            raise Bn254Error(-3 if variant == MIN_SIG else -2, "public-key buffer length")
        if agg_sig is not None and len(agg_sig) != rsig:
            raise Bn254Error(-2 if variant == MIN_SIG else -3, "signature length")
        part, bad = C.create_string_buffer(384), C.create_string_buffer(1)
        self._ck(self._L.bn254_aggregate_partial(self._h, variant, _ptr(pks), _ptr(blob), _ptr(off), n, _ptr(agg_sig), _ptr(dst), len(dst), flags,
                                                 _ptr(part), _ptr(bad)))
        return part.raw, bad.raw[0]

    def aggregate_finish(self, partials, bads):
        k = len(partials) // 384
        if k == 0 or len(partials) != 384 * k or len(bads) != k:
            raise Bn254Error(-4, "partials must be k x 384 bytes with k flags")
        ok = C.create_string_buffer(1)
        self._ck(self._L.bn254_aggregate_finish(self._h, _ptr(partials), _ptr(bytes(bads)), k, _ptr(ok)))
        return ok.raw[0]

    # ---- codecs, predicates, scalar multiplication ------------------------------------------------
    def _map(self, fn, data, rin, rout, status=False, extra=None):
        n = len(data) // rin
        if len(data) != rin * n:
            raise Bn254Error(-2, "length mismatch")
        out = C.create_string_buffer(max(1, rout * n))
        if extra is not None:
            if len(extra) != 32 * n:
                raise Bn254Error(-1, "scalars must be n x 32 bytes")
            self._ck(fn(self._h, _ptr(data), _ptr(extra), n, _ptr(out)))
            return out.raw[:rout * n]
        if status:
            st = C.create_string_buffer(max(1, n))
            self._ck(fn(self._h, _ptr(data), n, _ptr(out), _ptr(st)))
            return out.raw[:rout * n], st.raw[:n]
        self._ck(fn(self._h, _ptr(data), n, _ptr(out)))
        return out.raw[:rout * n]

    def g1_compress(self, pts): return self._map(self._L.bn254_g1_compress, pts, 64, 32)
    def g2_compress(self, pts): return self._map(self._L.bn254_g2_compress, pts, 128, 64)
    def g1_decompress(self, comp): return self._map(self._L.bn254_g1_decompress, comp, 32, 64, status=True)
    def g2_decompress(self, comp): return self._map(self._L.bn254_g2_decompress, comp, 64, 128, status=True)
    def g1_is_on_curve(self, pts): return self._map(self._L.bn254_g1_is_on_curve, pts, 64, 1)
    def g2_is_on_curve(self, pts): return self._map(self._L.bn254_g2_is_on_curve, pts, 128, 1)
    def g1_is_torsion_free(self, pts): return self._map(self._L.bn254_g1_is_torsion_free, pts, 64, 1)
    def g2_is_torsion_free(self, pts): return self._map(self._L.bn254_g2_is_torsion_free, pts, 128, 1)
    def g1_mul(self, pts, scalars): return self._map(self._L.bn254_g1_mul, pts, 64, 64, extra=scalars)
    def g2_mul(self, pts, scalars): return self._map(self._L.bn254_g2_mul, pts, 128, 128, extra=scalars)

    def g1_sum(self, pts):
        if len(pts) % 64:
            raise Bn254Error(-2, "length mismatch")
        out = C.create_string_buffer(64)
        self._ck(self._L.bn254_g1_sum(self._h, _ptr(pts), len(pts) // 64, _ptr(out)))
        return out.raw

    def g2_sum(self, pts):
        if len(pts) % 128:
            raise Bn254Error(-3, "length mismatch")
        out = C.create_string_buffer(128)
        self._ck(self._L.bn254_g2_sum(self._h, _ptr(pts), len(pts) // 128, _ptr(out)))
        return out.raw

    # ---- the C ABI itself, for callers that own their buffers ---------------------------------------
    def raw(self, name, *args):
        """Call the C entry point `name` (include/bls_bn254_b200.h) on this context with the arguments as given: integers
        are taken as addresses (e.g. the data_ptr() of a pinned torch tensor or of a numpy array), bytes / ctypes buffers /
        numpy arrays are passed by pointer.  This is the zero-copy form of every method above: no Python-side staging,
        results are written where the caller says.  Raises Bn254Error on a non-zero return code."""
        fn = getattr(self._L, name)
        conv = [a if isinstance(a, int) or a is None else _ptr(a) if isinstance(a, (bytes, bytearray)) or hasattr(a, "ctypes") else a for a in args]
        self._ck(fn(self._h, *conv))

    # ---- device-resident entry points (raw device addresses, e.g. torch.Tensor.data_ptr()) ---------
    def hash_to_g1_dev(self, msgs, offsets, n, dst, dst_len, count, out, status=None):
        self._ck(self._L.bn254_hash_to_g1_dev(self._h, msgs, offsets, n, dst, dst_len, count, out, status))

    def hash_to_g2_dev(self, msgs, offsets, n, dst, dst_len, count, out, status=None):
        self._ck(self._L.bn254_hash_to_g2_dev(self._h, msgs, offsets, n, dst, dst_len, count, out, status))

    def pairing_dev(self, g1, g2, n, gt, status=None):
        self._ck(self._L.bn254_pairing_batch_dev(self._h, g1, g2, n, gt, status))

    def verify_batch_dev(self, variant, pks, msgs, offsets, sigs, n, dst, dst_len, flags, ok):
        self._ck(self._L.bn254_verify_batch_dev(self._h, variant, pks, msgs, offsets, sigs, n, dst, dst_len, flags, ok))

    def verify_batch_randomized_dev(self, variant, pks, msgs, offsets, sigs, n, dst, dst_len, flags, seed, rand_bits, group, ok_groups):
        self._ck(self._L.bn254_verify_batch_randomized_dev(self._h, variant, pks, msgs, offsets, sigs, n, dst, dst_len, flags, seed, rand_bits, group,
                                                           ok_groups))

    def aggregate_partial_dev(self, variant, pks, msgs, offsets, n, agg_sig, dst, dst_len, flags, partial, bad):
        self._ck(self._L.bn254_aggregate_partial_dev(self._h, variant, pks, msgs, offsets, n, agg_sig, dst, dst_len, flags, partial, bad))

    def aggregate_verify_dist_dev(self, variant, pks, msgs, offsets, n, agg_sig, dst, dst_len, flags, ok):
        self._ck(self._L.bn254_aggregate_verify_dist_dev(self._h, variant, pks, msgs, offsets, n, agg_sig, dst, dst_len, flags, ok))

    def imad_peak(self, mode=2, iters=4096):
        """(IMAD-ops/s, ms); mode 0 mad.lo, 1 mad.lo+mad.hi, 2 IMAD.WIDE (2 ops each)."""
        ops, ms = C.c_double(), C.c_double()
        self._ck(self._L.bn254_imad_peak(self._h, int(mode), iters, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value


    def _fp_peak(self, mode, warps_per_sm, iters=4000):
        ops = C.c_double()
        self._ck(self._L.bn254_fp_peak(self._h, mode, warps_per_sm, iters, C.byref(ops)))
        return ops.value


Context.fp_peak = Context._fp_peak

# G1 / G2 generators in wire form (g1.rs:274-280; fp2.rs:305-348)
G1_GENERATOR = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
G2_GENERATOR = bytes.fromhex(
    "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"
    "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"
    "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"
    "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa")
DST_MIN_SIG = b"BLS_SIG_BN254G1_XMD:SHA-256_SVDW_RO_NUL_"
DST_MIN_PK = b"BLS_SIG_BN254G2_XMD:SHA-256_SVDW_RO_NUL_"
R_ORDER = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
