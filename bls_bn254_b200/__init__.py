"""bls_bn254_b200 -- B200-native batch BLS verification over BN254 (hash-to-curve, multi-Miller loop,
final exponentiation) behind a C ABI.  CUDA only: importing works anywhere, computing needs a GPU."""
from .api import (CHECK_ON_CURVE, CHECK_SUBGROUP, DST_MIN_PK, DST_MIN_SIG, G1_GENERATOR, G2_GENERATOR, MIN_PK, MIN_SIG, R_ORDER, Bn254Error, Context, PreparedKeys,  # noqa: F401
                  pack_msgs)
from . import scheme  # noqa: F401


def __getattr__(name):
    # bls_bn254_b200.dist needs torch.distributed: imported on first use only
    if name == "dist":
        import importlib
        return importlib.import_module(".dist", __name__)
    raise AttributeError(name)
