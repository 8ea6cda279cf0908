"""C++ host mirror of the reference optimizer interface (include/smpc_optimizer.hpp, OptimizerT<PlainMsgs>) over the
C-ABI: it must compile with a plain g++ (no CUDA headers, no ROS), fail loudly without a GPU, and on the GPU return
exactly what the Python host layer returns for the same scene (both marshal into smpc_optimize)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nav2_social_mpc_controller_b200")
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
YAML = os.path.join(PKG, "params", "soc_work_obst_in_benchmark.yaml")


def _build_cli():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "optimizer_cli")
    src = os.path.join(ROOT, "tests", "cpp", "optimizer_cli.cpp")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"), src,
           "-L" + PKG, "-lsmpc", "-Wl,-rpath," + PKG, "-Wl,--allow-shlib-undefined", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _env():
    """libsmpc.so needs libcudart.so.12: the copy torch ships, else the toolkit's."""
    dirs = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cuda_runtime as cr
        dirs.insert(0, os.path.join(list(cr.__path__)[0], "lib"))
    except Exception:
        pass
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(dirs + [env.get("LD_LIBRARY_PATH", "")])
    return env


def _write_scene(path, n_ticks, poses, cmds, people, speed, time_step, costmap, od):
    with open(path, "w") as f:
        def row(a):
            f.write(" ".join(repr(float(v)) for v in np.asarray(a, dtype=np.float64).ravel()) + "\n")
        f.write(f"{YAML} FollowPath {n_ticks}\n{len(poses)}\n")
        row(poses)
        f.write(f"{len(cmds)}\n")
        row(cmds)
        f.write(f"{len(people)}\n")
        row(people)
        f.write(f"{speed[0]!r} {speed[1]!r} {float(np.float32(time_step))!r}\n")
        f.write(f"{costmap.shape[1]} {costmap.shape[0]} 0.0 0.0 0.05\n")
        f.write(" ".join(str(int(v)) for v in costmap.ravel()) + "\n")
        f.write(f"{od['width']} {od['height']} {od['resolution']!r} {od['origin_x']!r} {od['origin_y']!r} "
                f"{len(od['indexes'])}\n")
        f.write(" ".join(repr(float(v)) for v in od["distances"]) + "\n")
        f.write(" ".join(str(int(v)) for v in od["indexes"]) + "\n")


def _parse(stdout):
    ticks = []
    lines = stdout.strip().splitlines()
    for i in range(0, len(lines), 4):
        head = lines[i].split()
        d = dict(ok=int(head[3]), termination=int(head[5]), iterations=int(head[7]), cost_initial=float(head[9]),
                 cost_final=float(head[11]))
        for key, width, line in (("path", 3, lines[i + 1]), ("cmds", 2, lines[i + 2]), ("proj", 18, lines[i + 3])):
            tok = line.split()
            assert tok[0] == key
            d[key] = np.array([float(v) for v in tok[2:]]).reshape(int(tok[1]), width)
        ticks.append(d)
    return ticks


def test_cpp_host_compiles_without_cuda_or_ros_headers():
    exe = _build_cli()
    assert os.path.exists(exe)
    # the header pair must also be self-contained
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), "-x", "c++", "-"],
                       input='#include "smpc_optimizer.hpp"\n#include "smpc_plain_msgs.hpp"\n'
                             "template class nav2_social_mpc_controller_b200::OptimizerT<"
                             "nav2_social_mpc_controller_b200::PlainMsgs>;\n",
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_host_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = _build_cli()
    scene = tmp_path / "scene.txt"
    cm = np.zeros((4, 4), dtype=np.uint8)
    od = dict(width=4, height=4, resolution=0.05, origin_x=0.0, origin_y=0.0, distances=np.zeros(16, np.float32),
              indexes=np.zeros(16, np.uint32))
    _write_scene(scene, 1, np.zeros((3, 3)), np.zeros((2, 2)), np.zeros((0, 5)), (0.0, 0.0), 0.05, cm, od)
    r = subprocess.run([exe, str(scene)], capture_output=True, text=True, env=_env())
    assert r.returncode == 1 and "exception" in r.stdout and "no CPU path" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("n_people", [0, 2, 5])
def test_cpp_optimizer_matches_python_host_layer(tmp_path, n_people):
    from nav2_social_mpc_controller_b200.optimizer import Optimizer, OptimizerParams
    from tests.test_gpu_optimize import _scene
    _, poses, cmds, people, speed, costmap, od = _scene("soc_work_obst", n_people, seed=11 + n_people)
    params = OptimizerParams.from_yaml(YAML, "FollowPath")
    exe = _build_cli()
    scene = tmp_path / "scene.txt"
    _write_scene(scene, 2, poses, cmds, people, speed, params.time_step, costmap, od)
    r = subprocess.run([exe, str(scene)], capture_output=True, text=True, env=_env())
    assert r.returncode == 0, r.stdout + r.stderr
    ticks = _parse(r.stdout)
    assert len(ticks) == 2
    opt = Optimizer(0)
    opt.initialize(params)
    try:
        for t in ticks:  # tick 2 exercises the per-handle warm-start memory on both sides
            ok, path, new_cmds, proj, info = opt.optimize(poses, cmds, people, speed, params.time_step, costmap,
                                                          (0.0, 0.0), 0.05, od)
            assert t["ok"] == int(ok)
            assert t["termination"] == info["termination"] and t["iterations"] == info["iterations"]
            assert t["cost_final"] == pytest.approx(info["cost_final"], rel=1e-8)
            assert np.allclose(t["proj"].reshape(-1, 3, 6), proj, rtol=0, atol=1e-9)
            if ok:
                # the C++ side carries yaw through a quaternion (ROS message shape): agreement to round-off, far
                # inside the 1e-6 band of the north star
                assert np.abs(t["cmds"] - new_cmds).max() <= 1e-6
                assert np.abs(t["path"][:, :2] - path[:, :2]).max() <= 1e-6
                assert np.abs(np.cos(t["path"][:, 2] - path[:, 2]) - 1).max() <= 1e-10
    finally:
        opt.close()


def test_ros_shim_compiles_against_stubs():
    """ros_shim/optimizer.hpp (the header a ROS 2 workspace swaps in for the reference's optimizer.hpp) type-checks
    against stand-in ROS headers and reproduces the defaults / the error of OptimizerParams::get
    (reference src/optimizer.cpp:26-84, :44)."""
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "ros_shim_check")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "ros_shim"), "-I" + os.path.join(ROOT, "tests", "cpp", "ros_stubs"),
           os.path.join(ROOT, "tests", "cpp", "ros_shim_check.cpp"), "-L" + PKG, "-lsmpc", "-Wl,-rpath," + PKG,
           "-Wl,--allow-shlib-undefined", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, env=_env())
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].split() == ["SPARSE_NORMAL_CHOLESKY", "1e-15", "1e-07", "1e-10", "100", "18", "5", "3", "90", "1.5"]
    assert lines[1] == "Invalid parameter: linear_solver_type"
