"""CPU-only tests of the host side: libsmpc.so loads and exports every symbol include/smpc.h declares, the
parameter defaults / yaml loader mirror OptimizerParams::get (reference src/optimizer.cpp:16-85), derived problem
sizes follow src/optimizer.cpp:248-249, and compute entries fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nav2_social_mpc_controller_b200 import _lib, abi, scenarios as sc
from nav2_social_mpc_controller_b200.optimizer import Optimizer, OptimizerParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARAMS = os.path.join(ROOT, "nav2_social_mpc_controller_b200", "params")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "smpc.h")).read()
    declared = set(re.findall(r"\b(smpc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/smpc.h but not exported by libsmpc.so"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert L.smpc_abi_version() == abi.SMPC_ABI_VERSION


def test_struct_layout_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "smpc.h")).read()

    def fields(struct):
        body = re.search(r"typedef struct " + struct + r" \{(.*?)\} " + struct + ";", hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return [re.sub(r"\[.*\]", "", d.strip().split()[-1].lstrip("*")) for d in body.split(";") if d.strip()]

    assert fields("smpc_params") == [f[0] for f in abi.SmpcParams._fields_]
    assert fields("smpc_batch") == [f[0] for f in abi.SmpcBatch._fields_]
    assert fields("smpc_result") == [f[0] for f in abi.SmpcResult._fields_]
    assert fields("smpc_eval_out") == [f[0] for f in abi.SmpcEvalOut._fields_]


def test_defaults_mirror_reference_declarations():
    p = OptimizerParams()
    # reference src/optimizer.cpp:26-84
    assert p.linear_solver_type == "SPARSE_NORMAL_CHOLESKY"
    assert (p.param_tol, p.fn_tol, p.gradient_tol, p.max_iterations) == (1e-15, 1e-7, 1e-10, 100)
    assert (p.control_horizon, p.parameter_block_length) == (5, 5)
    assert (p.distance_w, p.socialwork_w, p.velocity_w, p.angle_w, p.agent_angle_w) == (3.0, 1.0, 0.5, 0.0, 0.5)
    assert (p.proxemics_w, p.velocity_feasibility_w, p.obstacle_w, p.goal_align_w) == (90.0, 0.5, 0.0, 0.0)
    assert (p.current_path_w, p.current_cmds_w) == (1.0, 1.0)
    # reference src/path_trajectorizer.cpp:52-59, src/social_mpc_controller.cpp:59-65
    assert p.max_time == 3.0 and abs(p.time_step - 0.05) < 1e-8 and p.base_frame == "base_footprint"
    assert p.desired_linear_vel == 0.5 and abs(p.fov_angle - np.pi / 4) < 1e-15


@pytest.mark.parametrize("fname,name", [("params.yaml", "params_yaml"), ("obst_only_in_benchmark.yaml", "obst_only"),
                                        ("soc_work_obst_in_benchmark.yaml", "soc_work_obst"),
                                        ("readme_example.yaml", "readme")])
def test_yaml_loader_matches_parameter_sets(fname, name):
    p = OptimizerParams.from_yaml(os.path.join(PARAMS, fname), "FollowPath")
    want = sc.make_params(name)
    for f, _ in abi.SmpcParams._fields_:
        if f in ("desired_linear_vel", "fov_angle"):
            continue  # not part of the FollowPath subtree of the shipped yamls / scenario tables
        assert getattr(p.c, f) == getattr(want, f), f
    if name == "obst_only":
        assert p.proxemics_w == 90.0  # not set by the yaml -> default stays active


def test_yaml_loader_errors(tmp_path):
    bad = tmp_path / "bad.yaml"
    bad.write_text("FollowPath:\n  optimizer:\n    linear_solver_type: \"CGNR\"\n")
    with pytest.raises(_lib.SmpcError) as ei:
        OptimizerParams.from_yaml(str(bad), "FollowPath")
    assert "linear_solver_type" in str(ei.value)  # mirrors the std::runtime_error of src/optimizer.cpp:44
    with pytest.raises(_lib.SmpcError):
        OptimizerParams.from_yaml(str(bad), "NoSuchPlugin")
    with pytest.raises(_lib.SmpcError):
        OptimizerParams.from_yaml(str(tmp_path / "missing.yaml"), "FollowPath")


@pytest.mark.parametrize("ch,bl,S,want", [(18, 6, 13, (13, 6, 3, 2)), (18, 6, 28, (18, 6, 3, 3)),
                                          (20, 4, 38, (20, 4, 5, 5)), (18, 1, 28, (18, 1, 18, 18)),
                                          (5, 5, 3, (3, 3, 1, 1))])
def test_problem_dims(ch, bl, S, want):
    """SURVEY §8 size table; Q2: bounds cover ch/bl blocks while ceil(ch/bl) blocks exist."""
    p = abi.SmpcParams()
    _lib.lib().smpc_params_default(C.byref(p))
    p.control_horizon, p.parameter_block_length = ch, bl
    out = [C.c_int() for _ in range(4)]
    assert _lib.lib().smpc_problem_dims(C.byref(p), S, *[C.byref(o) for o in out]) == 0
    assert tuple(o.value for o in out) == want
    assert abi.problem_dims(ch, bl, S) == want


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    opt = Optimizer(0)
    with pytest.raises(_lib.SmpcError) as ei:
        opt.initialize(sc.make_params("obst_only"))
    assert "no CUDA device" in str(ei.value) or ei.value.code == -2
    with pytest.raises(RuntimeError):
        opt.solve_batch(sc.corridor(B=2))


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the package or the CUDA sources may reference it."""
    pkg = os.path.join(ROOT, "nav2_social_mpc_controller_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "liboracle" not in text and "oracle/" not in text.replace(
                    "under oracle/", ""), os.path.join(dirpath, f)


def test_scenario_generators_are_deterministic_and_well_formed():
    a, b = sc.corridor(B=16), sc.corridor(B=16)
    for k in a.arrays:
        if a.arrays[k] is not None:
            assert np.array_equal(a.arrays[k], b.arrays[k]), k
    c = sc.crowd(B=8, A=5)
    S, nb = c.n_steps, c.n_blocks
    assert c.arrays["agents"].shape == (8, 5, 6, S + 1)
    assert c.arrays["u0"].shape == (8, nb, 2) and c.arrays["path_xy"].shape == (8, 2, S + 1)
    # Q1: block b starts at the seed velocity of TIME INDEX b; index 0 is the measured speed
    assert np.all(c.arrays["u0"][:, 0, 0] <= 0.6)
    m = sc.multistart(n_robots=4, n_starts=8)
    assert m.n_problems == 32
    assert np.array_equal(m.arrays["u0"][0], sc.crowd(B=4, A=3, config_id=4, n_maps=4).arrays["u0"][0])
    assert np.all(m.arrays["u0"][..., 0] >= 0) and np.all(m.arrays["u0"][..., 0] <= 0.6)


def test_host_pipeline_chunk_plan():
    """Chunk plan of smpc_solve_batch (pure host logic, no GPU): people-free batches are one launch; batches with people
    are cut into <= 8 chunks of multiples of 256 problems, none smaller than the 16-warp threshold (2 * 16 * n_sm), and
    chunk starts keep b % M intact for shared costmaps without an index."""
    import ctypes as C
    from nav2_social_mpc_controller_b200 import _lib
    L = _lib.lib()
    L.smpc_debug_plan_chunks.restype = C.c_int
    L.smpc_debug_plan_chunks.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int)] * 2

    def plan(B, people, M, per_problem=0, index=0, forced=0, n_sm=148):
        n, c = C.c_int(), C.c_int()
        assert L.smpc_debug_plan_chunks(n_sm, B, people, M, per_problem, index, forced, C.byref(n), C.byref(c)) == 0
        return n.value, c.value

    assert plan(4096, 0, 4096, per_problem=1) == (1, 4096)          # BASELINE configs[1]: streamed, not chunked
    assert plan(65536, 0, 256) == (1, 65536)
    w16 = 2 * 16 * 148
    for B, M in [(65536, 256), (16384, 256), (262144, 256), (10000, 256), (9471, 1), (9999, 4), (1_000_000, 256)]:
        n, c = plan(B, 1, M)
        assert 1 <= n <= 8 and (n - 1) * c < B <= n * c
        if n > 1:
            assert c % 256 == 0 and c % M == 0                       # b % M unchanged at every chunk start
            assert B - (n - 1) * c >= w16 and c >= w16               # no chunk falls below the 16-warp threshold
    assert plan(65536, 1, 256)[0] == 8 and plan(16384, 1, 256)[0] == 3 and plan(9000, 1, 256)[0] == 1
    # shared maps without an index whose count does not divide any admissible chunk size: fall back to one launch
    assert plan(65536, 1, 1000) == (1, 65536)
    assert plan(65536, 1, 1000, index=1)[0] == 8                     # with an explicit index the maps need no alignment
    # the SMPC_CHUNKS override ignores the size rule but still honours the modulo rule
    assert plan(1300, 1, 256, forced=3) == (3, 512)
    assert L.smpc_debug_plan_chunks(0, 10, 0, 1, 0, 0, 0, None, None) != 0


def test_scenario_sharing_batches_expand_to_the_per_problem_layout():
    """smpc_batch.scenario_index (multi-start): the shared-scene form of a batch and its per-problem form describe the
    same problems; slices keep the scene arrays whole and cut only u0 / scenario_index."""
    plain = sc.multistart(n_robots=5, n_starts=7)
    shared = sc.multistart(n_robots=5, n_starts=7, shared=True)
    full = shared.expanded()
    assert shared.arrays["pose0"].shape[0] == 5 and shared.arrays["u0"].shape[0] == 35
    assert shared.struct().n_scenarios == 5 and plain.struct().n_scenarios == 0 and full.struct().n_scenarios == 0
    for k, v in plain.arrays.items():
        if v is not None:
            assert np.array_equal(v, full.arrays[k]), k
    part = shared.slice(7, 21)
    assert part.n_problems == 14 and part.arrays["pose0"].shape[0] == 5
    assert part.arrays["scenario_index"].tolist() == [1] * 7 + [2] * 7
    assert np.array_equal(part.expanded().arrays["agents"], plain.slice(7, 21).arrays["agents"])


def test_multi_gpu_shard_bounds_without_a_gpu():
    """smpc_debug_shard_bounds: the contiguous cut smpc_solve_batch_multi uses — shards cover [0, n) without gaps, start
    at multiples of the granule (multi-start: a robot's starts stay on one GPU), sizes differ by at most one granule."""
    import ctypes as C
    from nav2_social_mpc_controller_b200 import _lib
    L = _lib.lib()
    for n, world, g in [(4096, 8, 1), (262144, 8, 1024), (1000000, 3, 1), (100, 8, 16), (5, 8, 1), (1030, 4, 64)]:
        lo, hi = C.c_int(), C.c_int()
        prev, sizes = 0, []
        for r in range(world):
            assert L.smpc_debug_shard_bounds(n, world, r, g, C.byref(lo), C.byref(hi)) == 0
            assert lo.value == prev and hi.value >= lo.value
            assert lo.value % g == 0
            sizes.append(hi.value - lo.value)
            prev = hi.value
        assert prev == n
        full = [s for s in sizes[:-1]]
        assert max(full) - min(full) <= g
