"""A bounded slice of the randomised parity sweep (tests/fuzz_parity.py) inside the GPU suite: random sizes, margins, filter lists,
zones, rotation counts, angle / scale ranges, symmetry pools, the enhance step on or off — fused chain and batched operators
against the oracle (reference call sequences: overlays.py:104-149, pixels_isolés.py:29-61, rotations.py:96-101, tranfo.py:9-55).
Seeds 705192 (math.hypot vs libm hypot) and the pow(x, 2.0) case are the mismatches the sweep found in round 2."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fuzz():
    spec = importlib.util.spec_from_file_location("fuzz_parity", Path(__file__).with_name("fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("seed0", [3, 11])
def test_fuzz_slice(seed0):
    fz = _fuzz()
    for c in range(60):
        seed = seed0 * 100_000 + c
        fn = fz.chain_case if c % 2 == 0 else fz.ops_case
        fn(np.random.default_rng(seed), seed)


def test_fuzz_regression_seeds():
    fz = _fuzz()
    for seed in (705192,):
        fz.chain_case(np.random.default_rng(seed), seed)
