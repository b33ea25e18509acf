"""plume_b200 -- B200-native batch PLUME signer/verifier: Python host mirror of plume_rustcrypto's
API over the C ABI in include/plume_b200.h (CUDA kernels for sm_100a in ../csrc)."""
from .api import (DST, ORDER, PlumeContext, PlumeError, PlumeSignature, PlumeSignatureV1Fields, PlumeSigner,
                  SecretKey, default_context, encode_pt, hash_to_curve, pack_messages, point_from_bytes,
                  point_to_bytes, scalar_from_sec1_der, scalars_to_sec1_der)
from ._lib import LIB_PATH, SYMBOLS, load
from .shard import all_ranks_true, gather_counts, reduce_max, shard_range
from . import arkworks   # the twin crate's API (sign_with_r / sign / verify_non_zk), SURVEY.md 8f-3

__all__ = ["DST", "ORDER", "PlumeContext", "PlumeError", "PlumeSignature", "PlumeSignatureV1Fields", "PlumeSigner",
           "SecretKey", "default_context", "encode_pt", "hash_to_curve", "pack_messages", "point_from_bytes",
           "point_to_bytes", "scalar_from_sec1_der", "scalars_to_sec1_der", "arkworks", "LIB_PATH", "SYMBOLS", "load", "shard_range", "reduce_max", "all_ranks_true", "gather_counts"]
