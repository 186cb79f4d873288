"""Randomised parity sweep: task counts across all kernel families (1..32), ragged shapes (tiny, D < N, multi-tile),
mask strategies and densities, weightings, noise region, fp16 / fp32 bases, centring on / off, quantiser settings --
every case against the oracle on the same seeded inputs with the tolerances of tests/parity.py."""
import random

import pytest

from tests import parity

pytestmark = pytest.mark.gpu

TASK_COUNTS = [1, 2, 3, 4, 5, 6, 7, 8, 8, 8, 9, 10, 12, 14, 16, 17, 20, 20, 24, 25, 32]


def _case(rng):
    n = rng.choice(TASK_COUNTS)
    shapes = {}
    for i in range(rng.randint(2, 5)):
        kind = rng.choice(["mat", "vec", "tiny", "big"])
        if kind == "mat":
            shapes[f"p{i}.weight"] = (rng.randint(3, 300), rng.randint(2, 90))
        elif kind == "vec":
            shapes[f"p{i}.bias"] = (rng.randint(20, 5000),)
        elif kind == "tiny":
            shapes[f"p{i}.t"] = (rng.randint(1, 40),)
        else:
            shapes[f"p{i}.big"] = (rng.randint(20000, 60000),)
    mask_p = rng.choice([None, 0.2, 0.5, 0.9])
    kw = dict(svd_mask_strategy=rng.choice(["union", "intersection", "majority"]),
              svd_energy_threshold=rng.choice([0.5, 0.8, 0.9, 0.95, 0.99]),
              svd_low_bits=rng.choice([2, 4, 4, 8]), svd_rtvq_stages=rng.choice([1, 2, 2, 3]),
              svd_weighting=rng.choice(["uniform", "performance", "cluster"]),
              svd_fp16=rng.choice([True, True, False]), svd_center=rng.choice([True, True, False]),
              svd_include_noise=rng.choice([False, False, True]) and mask_p is not None,
              svd_noise_shrink=rng.choice([0.25, 0.5, 1.0]))
    if kw["svd_weighting"] == "cluster":
        kw["svd_cluster_k"] = 2 if n >= 2 else 1
    if kw["svd_weighting"] == "performance":
        kw["svd_weighting_temperature"] = rng.choice([1.0, 5.0])
    return n, shapes, mask_p, kw, rng.randint(0, 10 ** 6)


@pytest.mark.parametrize("block", range(8))
def test_random_configurations_against_oracle(cuda_device, block):
    failures = []
    for it in range(15):
        rng = random.Random(7000 + 100 * block + it)
        n, shapes, mask_p, kw, seed = _case(rng)
        try:
            ref, res, _ = parity.run_both(shapes, n, mask_p=mask_p, seed=seed, **kw)
            parity.compare_run(ref, res)
        except AssertionError as e:            # collect: one report with every failing configuration
            failures.append(f"n={n} shapes={shapes} mask_p={mask_p} seed={seed} {kw}: {str(e)[:200]}")
    assert not failures, "\n".join(failures)
