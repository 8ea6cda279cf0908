"""Golden vectors (tests/golden/solve_cases.npz): fixed level-1 inputs + the outputs the CPU oracle produced for them
when the fixture was generated. CPU test: today's oracle still reproduces them (guards the checker itself). GPU test:
the CUDA path, through the C-ABI, matches them within the north-star tolerances (controls 1e-6 absolute, final cost
1e-8 relative, same termination and iteration count)."""
import numpy as np
import pytest

from tests import golden_lib

CASES = golden_lib.load()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_vectors(oracle, name):
    batch, gold, ev = CASES[name]
    out = oracle.solve_batch(batch, want=tuple(gold))
    assert np.array_equal(out["termination"], gold["termination"])
    assert np.array_equal(out["iterations"], gold["iterations"])
    assert np.array_equal(out["usable"], gold["usable"])
    # same compiler flags (-ffp-contract=off) and libm on both sides: agreement to the last few bits
    assert np.abs(out["u"] - gold["u"]).max() <= 1e-12
    assert np.allclose(out["cost_final"], gold["cost_final"], rtol=1e-12, atol=0)
    assert np.abs(out["cmds"] - gold["cmds"]).max() <= 1e-12
    assert np.abs(out["path"][..., :2] - gold["path"][..., :2]).max() <= 1e-12
    for i in range(batch.n_problems):
        e = oracle.evaluate(batch, i, batch.arrays["u0"][i])
        assert e["cost"] == pytest.approx(ev["cost"][i], rel=1e-13)
        assert np.allclose(e["grad"], ev["grad"][i][:e["grad"].size], rtol=1e-11, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_matches_golden_vectors(name):
    from nav2_social_mpc_controller_b200.optimizer import Optimizer
    batch, gold, ev = CASES[name]
    opt = Optimizer(0)
    opt.initialize(batch.params)
    try:
        got = opt.solve_batch(batch, want=tuple(gold))
        e = opt.eval_batch(batch, batch.arrays["u0"])
    finally:
        opt.close()
    assert np.allclose(e["cost"], ev["cost"], rtol=1e-11, atol=0)
    assert np.allclose(e["grad"], ev["grad"], rtol=1e-9, atol=1e-9 * np.abs(ev["grad"]).max())
    assert np.array_equal(got["usable"], gold["usable"])
    assert np.array_equal(got["termination"], gold["termination"])
    assert np.array_equal(got["iterations"], gold["iterations"])
    assert np.abs(got["u"] - gold["u"]).max() <= 1e-6                      # north star: controls, absolute
    assert np.abs(got["cmds"] - gold["cmds"]).max() <= 1e-6
    rel = np.abs(got["cost_final"] - gold["cost_final"]) / np.abs(gold["cost_final"])
    assert rel.max() <= 1e-8                                               # north star: final cost, relative
    assert np.abs(got["path"][..., :2] - gold["path"][..., :2]).max() <= 1e-6
