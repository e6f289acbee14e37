//! Error type of the GPU back end: the reference's four variants (error.rs:5-10) plus the conditions a device adds.
use core::fmt;

#[derive(Debug, Clone, PartialEq, Eq)]
pub enum Bn254Error {
    InvalidScalarBytes,
    InvalidG1Bytes,
    InvalidG2Bytes,
    InvalidGtBytes,
    /// a malformed call (null pointer, lengths that do not match, unknown variant)
    BadArgument(String),
    /// no usable CUDA device: the library has NO CPU fallback
    NoDevice,
    Cuda(i32, String),
    Nccl(i32, String),
}

impl Bn254Error {
    pub(crate) fn from_code(rc: i32, msg: String) -> Self {
        match rc {
            -1 => Self::InvalidScalarBytes,
            -2 => Self::InvalidG1Bytes,
            -3 => Self::InvalidG2Bytes,
            -4 => Self::InvalidGtBytes,
            -10 => Self::BadArgument(msg),
            -20 => Self::NoDevice,
            rc if rc <= -2000 => Self::Nccl(-2000 - rc, msg),
            rc => Self::Cuda(-100 - rc, msg),
        }
    }
}

impl fmt::Display for Bn254Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            Self::InvalidScalarBytes => write!(f, "invalid scalar bytes"),
            Self::InvalidG1Bytes => write!(f, "invalid G1 bytes"),
            Self::InvalidG2Bytes => write!(f, "invalid G2 bytes"),
            Self::InvalidGtBytes => write!(f, "invalid Gt bytes"),
            Self::BadArgument(m) => write!(f, "bad argument: {m}"),
            Self::NoDevice => write!(f, "no usable CUDA device (bls_bn254_b200 has no CPU path)"),
            Self::Cuda(k, m) => write!(f, "CUDA error {k}: {m}"),
            Self::Nccl(k, m) => write!(f, "NCCL error {k}: {m}"),
        }
    }
}
impl std::error::Error for Bn254Error {}
