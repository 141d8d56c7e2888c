"""The oracle against the REAL reference (JohnVithor/RL-Rust with the injected Philox stream, oracle/rust_ref/).
Runs only where oracle/_ref/parity_dump has been built (oracle/rust_ref/build_ref.sh — needs a Rust toolchain, which
this image lacks): skipped otherwise, and the parity claim stays "unpinned by the reference"."""
import json
import os
import subprocess

import numpy as np
import pytest

import parity as P
from oracle import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "parity_dump")
pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/parity_dump not built (no Rust toolchain here)")

CELLS = [dict(env=e, agent=a, selector=s, policy=p, target=t, real=1)
         for e in (0, 1, 2, 3) for a in (0, 1) for s in (0, 1) for p in (0, 1) for t in (0, 1, 2)]


@pytest.mark.parametrize("c", CELLS, ids=P.combo_id)
def test_oracle_equals_the_reference(c):
    n_ep, eval_at, seed, agent_id = 60, 6, 0x5EED0001, 3
    out = subprocess.run([EXE, str(c["env"]), str(c["agent"]), str(c["selector"]), str(c["policy"]), str(c["target"]), hex(seed),
                          str(agent_id), str(n_ep), str(eval_at), "100", "1", "8x8"], capture_output=True, text=True, check=True).stdout
    ref = json.loads(out)
    s = O.Session(P.oracle_config(c, P.hyper(n_ep, seed=seed)), agent_id)
    ret, ln, _, _ = s.train(n_ep, eval_at)
    te = s.training_error()
    st = s.export()[2]
    s.close()
    assert list(ln) == ref["episode_length"]
    assert list(ret.view(np.uint64)) == ref["reward_history_bits"]
    got, want = te.view(np.uint64), np.array(ref["training_error_bits"], np.uint64)
    nan = np.isnan(te) & np.isnan(want.view(np.float64))
    assert got.shape == want.shape and bool(np.all((got == want) | nan))
    assert st.rng_n == ref["rng_words"]
