"""An independent pin of the RNG contract: NVIDIA's own Philox4x32-10 (cuRAND device API, compiled
from tests/cuda/philox_curand.cu by __graft_entry__.build()) produces the same words as the engine's per-agent stream —
both the raw block function and curand_init(seed, subsequence = agent, offset = word) + curand()."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cuda", "philox_curand")


@pytest.mark.parametrize("seed,agent,first", [(0x5EED0001, 0, 0), (0xDEADBEEFCAFEF00D, 123456789012, 4096), (1, 2**40 + 7, 2**34), (0, 0xFFFFFFFF, 8)])
def test_stream_equals_curand_philox(rlb, seed, agent, first):
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    n = 1024
    out = subprocess.run([EXE, hex(seed), str(agent), str(first), str(n)], capture_output=True, text=True, check=True).stdout.split()
    raw = np.array([int(x, 16) for x in out[0::2]], np.uint32)
    api = np.array([int(x, 16) for x in out[1::2]], np.uint32)
    ours = rlb.abi.rng_words(seed, agent, first, n)
    assert np.array_equal(raw, ours)
    assert np.array_equal(api, ours)
