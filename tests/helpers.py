"""Shared helpers of the test-suite (oracle <-> product name mapping, bit comparison)."""
import json

import numpy as np

QT_NAMES = ("int4", "uint4", "int8", "uint8")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def as_i8(a, qt):
    """ml_dtypes / numpy quantized array → plain int8/uint8 values for comparison with fixtures."""
    return np.asarray(a).astype(np.int8 if qt.startswith("int") else np.uint8)


def golden_keys(npz):
    return json.loads(str(npz["keys"]))


def parse_rtn_key(key):
    wname, qt, strategy, gs, sym, rr, clip, mse = key.split("|")
    return wname, qt, strategy, int(gs), bool(int(sym)), bool(int(rr)), float(clip), bool(int(mse))


def stable_seed(*parts) -> int:
    """A seed that depends only on the VALUES of ``parts`` (CRC-32 of their repr) — unlike
    ``hash()``, whose string hashing is salted per process, so a failing case can be replayed."""
    import zlib
    return zlib.crc32(repr(parts).encode())
