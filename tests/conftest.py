import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure), built on demand with gcc."""
    import oracle
    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ffi():
    """The product's C-ABI library through ctypes; built on demand with nvcc if missing."""
    from halo2_aes_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    _ffi.lib()
    return _ffi


SEED = 0x68326165  # "h2ae"


def splitmix_bytes(seed: int, n: int) -> np.ndarray:
    """Counter-based SplitMix64 byte stream (the synthetic-input generator of SURVEY 8d)."""
    idx = np.arange((n + 7) // 8, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.view(np.uint8)[:n].copy()


R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def seeded_fr(seed: int) -> np.ndarray:
    """A uniform element of Fr (512 pseudo-random bits mod r) as Montgomery limbs."""
    b = splitmix_bytes(seed, 64)
    v = int.from_bytes(b.tobytes(), "little") % R_MOD
    m = v * ((1 << 256) % R_MOD) % R_MOD
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


@pytest.fixture(scope="session")
def challenges():
    return seeded_fr(SEED + 1), seeded_fr(SEED + 2), seeded_fr(SEED + 3)
