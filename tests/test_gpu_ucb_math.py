"""The hand-scheduled division / square root of the UCB bonus (rlb_device.cuh div_fast / sqrt_fast) against the
compiler's own `/` and `sqrt` — correctly rounded IEEE operations — bit for bit over the operands the bonus can see:
a = ln t, b = n (upper_confidence_bound.rs:33-37)."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("t_max,n_max,samples", [(64, 64, 1 << 22),             # every small pair many times over
                                                  (1 << 20, 1 << 20, 1 << 30),    # the range a bench run reaches
                                                  (1 << 40, (1 << 32) - 1, 1 << 30)])
def test_fast_div_sqrt_equal_the_compilers(rlb, t_max, n_max, samples):
    bad = C.c_uint64(123)
    for seed in (0, 0x5EED):
        rlb.abi.check(rlb.abi.lib.rlb_selftest_ucb_math(0, samples, t_max, n_max, seed, C.byref(bad)))
        assert bad.value == 0
