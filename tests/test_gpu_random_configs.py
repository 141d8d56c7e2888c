"""GPU parity on randomly drawn hyper-parameters (seeded): multiplicative epsilon decay, non-zero / negative / negative-zero
default Q, negative gamma*lambda (eligibilities of both signs, -0.0 products), large learning rates (overflow to inf/NaN is
legal and must match), odd max_steps, final_epsilon above the decayed value, other seeds and first agent ids — on every
table store the configuration supports."""
import numpy as np
import pytest

import parity as P
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def draw(rng):
    env = int(rng.integers(0, 4))
    c = dict(env=env, agent=int(rng.integers(0, 2)), selector=int(rng.integers(0, 2)), policy=int(rng.integers(0, 2)),
             target=int(rng.integers(0, 3)), real=int(rng.integers(0, 2)))
    n_ep = int(rng.integers(3, 16))
    h = P.hyper(n_ep,
                map_id=int(rng.integers(0, 2)), slippery=bool(rng.integers(0, 2)), max_steps=int(rng.choice([1, 3, 17, 64, 100, 250])),
                lr=float(rng.choice([0.05, 0.5, 1.0, 1.5, 1e-3, 7.0])), gamma=float(rng.choice([0.95, 1.0, 0.0, -0.7, 0.5])),
                lambda_=float(rng.choice([0.5, 1.0, 0.0, -0.9, 0.99])), eps0=float(rng.choice([1.0, 0.3, 0.0, 0.999])),
                eps_final=float(rng.choice([0.0, 0.05, 0.5])), ucb_c=float(rng.choice([0.5, 0.0, 2.0, -1.0])),
                default_q=float(rng.choice([0.0, -0.0, 1.0, -3.5, 100.0])), decay_kind=int(rng.integers(0, 2)),
                seed=int(rng.integers(0, 2 ** 63)))
    h["eps_decay"] = float(rng.choice([0.9, 0.5, 0.999])) if h["decay_kind"] == 1 else float(rng.choice([1.0 / (0.5 * n_ep), 0.01, 0.3]))
    return c, h, n_ep, int(rng.integers(2, 9)), int(rng.choice([1, 7, 32, 33, 40])), int(rng.integers(0, 2 ** 40))


@pytest.mark.parametrize("case", range(48))
def test_random_configuration(case):
    rng = np.random.default_rng(1000 + case)
    c, h, n_ep, eval_at, n_agents, first = draw(rng)
    o = O.batch_train(P.oracle_config(c, h), first, n_agents, n_ep, eval_at, n_threads=8)
    for store in ((1, 2, 3) if c["env"] in (1, 2) else (1,)) + ((4,) if c["agent"] == 1 else ()):
        try:
            g = P.gpu_run(c, h, n_agents, n_ep, eval_at, first_agent_id=first, store_kind=store)
        except Exception as exc:   # noqa: BLE001
            if store != 1 and getattr(exc, "status", None) == 5:   # this configuration does not fit that store
                continue
            raise
        P.compare(g, o, c)
