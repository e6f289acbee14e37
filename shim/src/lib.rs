//! Safe Rust over `libbls_bn254_b200.so`: the batch forms of the reference crate's hot-path functions, on the reference's
//! own byte formats (64 B G1, 128 B G2 as c1||c0, 384 B Gt; `include/bls_bn254_b200.h`).
//!
//! Reference function (mikelodder7/bls-bn254, src/inner_types)  ->  here
//!   G1Projective::hash / ::encode  (g1.rs:910-928)              Context::hash_to_g1 / encode_to_g1
//!   G2Projective::hash / ::encode  (g2.rs:919-936)              Context::hash_to_g2 / encode_to_g2
//!   pairing                        (pairings.rs:760-802)        Context::pairing
//!   multi_miller_loop              (pairings.rs:808-857)        Context::multi_miller_loop
//!   MillerLoopResult::final_exponentiation (50-178)             Context::final_exponentiation
//!   verify / aggregate_verify (absent upstream) Context::verify_batch / aggregate_verify
//!
//! A crate that wants the single-item signatures of the reference calls these with one element (see `hash_g1_one`,
//! `pairing_one`); throughput needs batches (a lone pairing on a GPU takes ~5 ms, 2^16 of them 24 ms).
pub mod error;
pub mod ffi;

pub use error::Bn254Error;
use std::ffi::CStr;

pub const MIN_SIG: i32 = ffi::BN254_MIN_SIG;
pub const MIN_PK: i32 = ffi::BN254_MIN_PK;
pub const CHECK_ON_CURVE: u32 = ffi::BN254_FLAG_CHECK_ON_CURVE;
pub const CHECK_SUBGROUP: u32 = ffi::BN254_FLAG_CHECK_SUBGROUP;

/// `bn254_ctx`: streams and scratch memory on one or more GPUs.  Single caller at a time (`&mut self`).
pub struct Context {
    h: *mut ffi::Bn254Ctx,
}
unsafe impl Send for Context {}

/// message i = blob[offsets[i] .. offsets[i + 1]]
pub fn pack_msgs<M: AsRef<[u8]>>(msgs: &[M]) -> (Vec<u8>, Vec<u64>) {
    let mut blob = Vec::new();
    let mut off = Vec::with_capacity(msgs.len() + 1);
    for m in msgs {
        off.push(blob.len() as u64);
        blob.extend_from_slice(m.as_ref());
    }
    off.push(blob.len() as u64);
    (blob, off)
}

impl Context {
    /// devices = CUDA device indices; several devices shard every batch by contiguous index range.
    pub fn new(devices: &[i32]) -> Result<Self, Bn254Error> {
        let mut h = core::ptr::null_mut();
        let rc = unsafe { ffi::bn254_ctx_create(&mut h, devices.as_ptr(), devices.len() as i32) };
        if rc != 0 {
            return Err(Bn254Error::from_code(rc, "bn254_ctx_create".into()));
        }
        Ok(Self { h })
    }

    fn ck(&self, rc: i32) -> Result<(), Bn254Error> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(ffi::bn254_last_error(self.h)) }.to_string_lossy().into_owned();
        Err(Bn254Error::from_code(rc, msg))
    }

    fn hash(&mut self, f: unsafe extern "C" fn(*mut ffi::Bn254Ctx, *const u8, *const u64, usize, *const u8, usize, *mut u8, *mut u8) -> i32,
            rec: usize, blob: &[u8], offsets: &[u64], dst: &[u8]) -> Result<(Vec<u8>, Vec<u8>), Bn254Error> {
        let n = offsets.len().saturating_sub(1);
        if offsets.last().map_or(false, |&e| e as usize > blob.len()) {
            return Err(Bn254Error::BadArgument("offsets[n] exceeds the message blob".into()));
        }
        let (mut out, mut st) = (vec![0u8; rec * n], vec![0u8; n]);
        let rc = unsafe { f(self.h, blob.as_ptr(), offsets.as_ptr(), n, dst.as_ptr(), dst.len(), out.as_mut_ptr(), st.as_mut_ptr()) };
        self.ck(rc)?;
        Ok((out, st))
    }
    /// G1Projective::hash (g1.rs:910-919) for every message: n x 64 B uncompressed points + one status byte each
    /// (0 = SvdW exceptional input, where the reference panics: fp.rs:325).
    pub fn hash_to_g1(&mut self, blob: &[u8], offsets: &[u64], dst: &[u8]) -> Result<(Vec<u8>, Vec<u8>), Bn254Error> {
        self.hash(ffi::bn254_hash_to_g1, 64, blob, offsets, dst)
    }
    pub fn encode_to_g1(&mut self, blob: &[u8], offsets: &[u64], dst: &[u8]) -> Result<(Vec<u8>, Vec<u8>), Bn254Error> {
        self.hash(ffi::bn254_encode_to_g1, 64, blob, offsets, dst)
    }
    pub fn hash_to_g2(&mut self, blob: &[u8], offsets: &[u64], dst: &[u8]) -> Result<(Vec<u8>, Vec<u8>), Bn254Error> {
        self.hash(ffi::bn254_hash_to_g2, 128, blob, offsets, dst)
    }
    pub fn encode_to_g2(&mut self, blob: &[u8], offsets: &[u64], dst: &[u8]) -> Result<(Vec<u8>, Vec<u8>), Bn254Error> {
        self.hash(ffi::bn254_encode_to_g2, 128, blob, offsets, dst)
    }

    /// pairing() (pairings.rs:760-802) + Gt::to_repr (499-514) for n pairs: n x 384 B and a status byte per pair
    pub fn pairing(&mut self, g1: &[u8], g2: &[u8]) -> Result<(Vec<u8>, Vec<u8>), Bn254Error> {
        let n = g1.len() / 64;
        if g1.len() != 64 * n {
            return Err(Bn254Error::InvalidG1Bytes);
        }
        if g2.len() != 128 * n {
            return Err(Bn254Error::InvalidG2Bytes);
        }
        let (mut gt, mut st) = (vec![0u8; 384 * n], vec![0u8; n]);
        let rc = unsafe { ffi::bn254_pairing_batch(self.h, g1.as_ptr(), g2.as_ptr(), n, gt.as_mut_ptr(), st.as_mut_ptr()) };
        self.ck(rc)?;
        Ok((gt, st))
    }
    /// multi_miller_loop (pairings.rs:808-857): ONE un-exponentiated MillerLoopResult (384 B) for n pairs
    pub fn multi_miller_loop(&mut self, g1: &[u8], g2: &[u8]) -> Result<[u8; 384], Bn254Error> {
        let n = g1.len() / 64;
        if g1.len() != 64 * n {
            return Err(Bn254Error::InvalidG1Bytes);
        }
        if g2.len() != 128 * n {
            return Err(Bn254Error::InvalidG2Bytes);
        }
        let mut ml = [0u8; 384];
        let rc = unsafe { ffi::bn254_multi_miller_loop(self.h, g1.as_ptr(), g2.as_ptr(), n, ml.as_mut_ptr()) };
        self.ck(rc)?;
        Ok(ml)
    }
    /// MillerLoopResult::final_exponentiation (pairings.rs:50-178) for n results
    pub fn final_exponentiation(&mut self, ml: &[u8]) -> Result<(Vec<u8>, Vec<u8>), Bn254Error> {
        let n = ml.len() / 384;
        if ml.len() != 384 * n {
            return Err(Bn254Error::InvalidGtBytes);
        }
        let (mut gt, mut st) = (vec![0u8; 384 * n], vec![0u8; n]);
        let rc = unsafe { ffi::bn254_final_exponentiation(self.h, ml.as_ptr(), n, gt.as_mut_ptr(), st.as_mut_ptr()) };
        self.ck(rc)?;
        Ok((gt, st))
    }

    fn recs(variant: i32) -> (usize, usize) {
        if variant == MIN_SIG { (128, 64) } else { (64, 128) }
    }
    /// n accept bytes (0 / 1): min-sig e(sig, -g2) e(H(m), pk) == 1, min-pk e(-g1, sig) e(pk, H(m)) == 1.
    /// Identity keys / signatures never verify; `flags` adds on-curve / subgroup validation of untrusted inputs.
    pub fn verify_batch(&mut self, variant: i32, pks: &[u8], blob: &[u8], offsets: &[u64], sigs: &[u8], dst: &[u8], flags: u32)
                        -> Result<Vec<u8>, Bn254Error> {
        let n = offsets.len().saturating_sub(1);
        let (rpk, rsig) = Self::recs(variant);
        if pks.len() != rpk * n {
            return Err(if variant == MIN_SIG { Bn254Error::InvalidG2Bytes } else { Bn254Error::InvalidG1Bytes });
        }
        if sigs.len() != rsig * n {
            return Err(if variant == MIN_SIG { Bn254Error::InvalidG1Bytes } else { Bn254Error::InvalidG2Bytes });
        }
        if offsets.last().map_or(false, |&e| e as usize > blob.len()) {
            return Err(Bn254Error::BadArgument("offsets[n] exceeds the message blob".into()));
        }
        let mut ok = vec![0u8; n];
        let rc = unsafe {
            ffi::bn254_verify_batch(self.h, variant, pks.as_ptr(), blob.as_ptr(), offsets.as_ptr(), sigs.as_ptr(), n, dst.as_ptr(), dst.len(), flags,
                                    ok.as_mut_ptr())
        };
        self.ck(rc)?;
        Ok(ok)
    }
    /// One accept byte per group of `group` consecutive tuples (small-exponent test; `seed` = 32 fresh CSPRNG bytes per call).
    #[allow(clippy::too_many_arguments)]
    pub fn verify_batch_randomized(&mut self, variant: i32, pks: &[u8], blob: &[u8], offsets: &[u64], sigs: &[u8], dst: &[u8], flags: u32,
                                   seed: &[u8; 32], rand_bits: i32, group: usize) -> Result<Vec<u8>, Bn254Error> {
        let n = offsets.len().saturating_sub(1);
        let (rpk, rsig) = Self::recs(variant);
        if pks.len() != rpk * n || sigs.len() != rsig * n || group == 0 {
            return Err(Bn254Error::BadArgument("buffer lengths / group".into()));
        }
        let mut ok = vec![0u8; (n + group - 1) / group];
        let rc = unsafe {
            ffi::bn254_verify_batch_randomized(self.h, variant, pks.as_ptr(), blob.as_ptr(), offsets.as_ptr(), sigs.as_ptr(), n, dst.as_ptr(), dst.len(),
                                               flags, seed.as_ptr(), rand_bits, group, ok.as_mut_ptr())
        };
        self.ck(rc)?;
        Ok(ok)
    }
    /// Same-key randomised verification: key k owns tuples `key_offsets[k] .. key_offsets[k + 1]` of `blob` / `sigs` (tuples
    /// sorted by key); one accept byte per KEY from one pairing check over the randomised sums of its signatures and message
    /// points (`seed` = 32 fresh CSPRNG bytes per call).
    #[allow(clippy::too_many_arguments)]
    pub fn verify_same_key_randomized(&mut self, variant: i32, pks: &[u8], key_offsets: &[u64], blob: &[u8], offsets: &[u64], sigs: &[u8],
                                      dst: &[u8], flags: u32, seed: &[u8; 32], rand_bits: i32) -> Result<Vec<u8>, Bn254Error> {
        let nk = key_offsets.len().saturating_sub(1);
        let n = offsets.len().saturating_sub(1);
        let (rpk, rsig) = Self::recs(variant);
        if pks.len() != rpk * nk || sigs.len() != rsig * n || key_offsets.first() != Some(&0) || key_offsets.last() != Some(&(n as u64)) {
            return Err(Bn254Error::BadArgument("buffer lengths / key offsets".into()));
        }
        let mut ok = vec![0u8; nk];
        let rc = unsafe {
            ffi::bn254_verify_same_key_randomized(self.h, variant, pks.as_ptr(), key_offsets.as_ptr(), nk, blob.as_ptr(), offsets.as_ptr(), sigs.as_ptr(),
                                                  dst.as_ptr(), dst.len(), flags, seed.as_ptr(), rand_bits, ok.as_mut_ptr())
        };
        self.ck(rc)?;
        Ok(ok)
    }
    /// Distinct-message aggregate verification: one multi-Miller product, one final exponentiation.
    pub fn aggregate_verify(&mut self, variant: i32, pks: &[u8], blob: &[u8], offsets: &[u64], agg_sig: &[u8], dst: &[u8], flags: u32)
                            -> Result<bool, Bn254Error> {
        let n = offsets.len().saturating_sub(1);
        let (rpk, rsig) = Self::recs(variant);
        if pks.len() != rpk * n || agg_sig.len() != rsig {
            return Err(Bn254Error::BadArgument("buffer lengths".into()));
        }
        let mut ok = 0u8;
        let rc = unsafe {
            ffi::bn254_aggregate_verify(self.h, variant, pks.as_ptr(), blob.as_ptr(), offsets.as_ptr(), n, agg_sig.as_ptr(), dst.as_ptr(), dst.len(), flags,
                                        &mut ok)
        };
        self.ck(rc)?;
        Ok(ok == 1)
    }

    // ---- the reference's single-item shapes ----------------------------------------------------------------------
    /// `G1Projective::hash::<ExpandMsgXmd<Sha256>>(msg, dst)` as 64 uncompressed bytes (feed to `G1Affine::from_uncompressed`, g1.rs:336-358)
    pub fn hash_g1_one(&mut self, msg: &[u8], dst: &[u8]) -> Result<[u8; 64], Bn254Error> {
        let (out, st) = self.hash_to_g1(msg, &[0, msg.len() as u64], dst)?;
        assert_eq!(st[0], 1, "SvdW exceptional input (the reference panics here too: fp.rs:325)");
        Ok(out.try_into().unwrap())
    }
    /// `pairing(&p, &q)` on uncompressed bytes -> `Gt::to_repr()` bytes (feed to `Gt::from_repr`, pairings.rs:516-579)
    pub fn pairing_one(&mut self, p: &[u8; 64], q: &[u8; 128]) -> Result<[u8; 384], Bn254Error> {
        let (gt, st) = self.pairing(p, q)?;
        if st[0] != 1 {
            return Err(Bn254Error::InvalidG1Bytes);
        }
        Ok(gt.try_into().unwrap())
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::bn254_ctx_destroy(self.h) }
    }
}
