"""Fleet tick on one GPU: the whole Optimizer::optimize pipeline (reference src/optimizer.cpp:148-452) for B robots
per call, every stage a CUDA kernel of libsmpc.so — people_to_status, format_to_optimize (+ per-robot warm-start
memory instead of the TrajectoryMemory singleton), project_people (SFM), the bounded TR-LM solve with its post-solve
expansion, and the memory update. torch is used only to own the device buffers and the stream."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, abi
from .optimizer import Optimizer


class FleetOptimizer:
    def __init__(self, params, n_robots: int, n_agents: int = 3, device: int = 0):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.opt = Optimizer(device)
        self.opt.initialize(params)
        self.p = self.opt.params
        self.B, self.A = n_robots, n_agents
        self.prev_poses = None  # [B][n][3] previous optimised path (device)
        self.prev_cmds = None   # [B][n][2]
        self.stream = torch.cuda.Stream(device=self.dev)

    def close(self):
        self.opt.close()

    def reset_memory(self):
        self.prev_poses = self.prev_cmds = None

    def _t(self, a, dtype):
        return self.torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).to(self.dev)

    def trajectorize_batch(self, global_path, pose):
        """Batched PathTrajectorizer::trajectorize (reference src/path_trajectorizer.cpp:120-288) on the GPU.
        global_path [B][N][2], pose [B][3] -> poses [B][max_steps+1][3], cmds [B][max_steps][3], n_steps [B]."""
        torch, L, h, p = self.torch, _lib.lib(), self.opt._h, self.p
        gp = np.ascontiguousarray(global_path, dtype=np.float64)
        B, N, _ = gp.shape
        max_steps = int(round(float(p.max_time) / round(float(p.time_step), 6)))
        with torch.cuda.stream(self.stream):
            d_gp, d_pose = self._t(gp, np.float64), self._t(pose, np.float64)
            poses = torch.zeros(B, max_steps + 1, 3, dtype=torch.float64, device=self.dev)
            cmds = torch.zeros(B, max_steps, 3, dtype=torch.float64, device=self.dev)
            n_steps = torch.zeros(B, dtype=torch.int32, device=self.dev)
            a = abi.SmpcTrajectorizeArgs()
            a.n_problems, a.n_path, a.max_steps, a.omnidirectional = B, N, max_steps, int(p.omnidirectional)
            a.desired_linear_vel, a.lookahead_dist = p.traj_desired_linear_vel, p.lookahead_dist
            a.max_angular_vel, a.time_step = p.max_angular_vel, round(float(p.time_step), 6)
            a.global_path, a.path_index, a.pose = d_gp.data_ptr(), None, d_pose.data_ptr()
            a.poses, a.cmds, a.n_steps = poses.data_ptr(), cmds.data_ptr(), n_steps.data_ptr()
            _lib.check(L.smpc_trajectorize_batch_device(h, C.byref(a), self.stream.cuda_stream))
        self.stream.synchronize()
        return poses.cpu().numpy(), cmds.cpu().numpy(), n_steps.cpu().numpy()

    def optimize_batch(self, poses, cmds, people_raw, n_people, speed, costmaps, costmap_origin, costmap_resolution,
                       od: dict, costmap_index=None, od_index=None) -> dict:
        """poses [B][n][3], cmds [B][n-1][2] (trajectorizer seeds, same length for the fleet), people_raw [B][A][5],
        n_people [B], speed [B][2], costmaps [M][sy][sx] u8, od = dict(width, height, resolution, origins [Mo][2],
        indexes u32 [Mo][h*w]). Returns host numpy: optimized [B] (the reference's bool), cmds [B][n'][2],
        path [B][n'][3], people_proj [B][A][6][n'], termination, iterations, cost_final."""
        torch, L, h = self.torch, _lib.lib(), self.opt._h
        p = self.p
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        cmds = np.ascontiguousarray(cmds, dtype=np.float64)
        B, n_in, _ = poses.shape
        assert B == self.B
        maxsize = int(round(float(np.float32(p.max_time) / np.float32(p.time_step))))
        n = n_in if n_in <= maxsize else maxsize - 1  # the cut of src/optimizer.cpp:492-497
        poses, cmds = poses[:, :n], cmds[:, : n - 1]
        S = n - 1
        ch, bl, nb, _ = self.opt.dims(S)
        st = self.stream.cuda_stream
        f64 = torch.float64
        with torch.cuda.stream(self.stream):
            d_poses, d_cmds = self._t(poses, np.float64), self._t(cmds, np.float64)
            d_speed = self._t(speed, np.float64)
            d_raw, d_np = self._t(people_raw, np.float64), self._t(n_people, np.int32)
            d_maps, d_morg = self._t(costmaps, np.uint8), self._t(costmap_origin, np.float64)
            d_midx = None if costmap_index is None else self._t(costmap_index, np.int32)
            d_oorg = self._t(np.asarray(od["origins"], dtype=np.float64).reshape(-1, 2), np.float64)
            d_oidx = self._t(np.asarray(od["indexes"], dtype=np.uint32).reshape(d_oorg.shape[0], -1).view(np.int32),
                             np.int32)
            d_osel = None if od_index is None else self._t(od_index, np.int32)
            init = torch.empty(B, self.A, 6, dtype=f64, device=self.dev)
            has_people = torch.empty(B, dtype=torch.uint8, device=self.dev)
            robot = torch.empty(B, n, 6, dtype=f64, device=self.dev)
            pose0 = torch.empty(B, 3, dtype=f64, device=self.dev)
            u0 = torch.empty(B, nb, 2, dtype=f64, device=self.dev)
            path_xy = torch.empty(B, 2, n, dtype=f64, device=self.dev)
            goal_yaw = torch.empty(B, dtype=f64, device=self.dev)
            agents = torch.empty(B, self.A, 6, n, dtype=f64, device=self.dev)
            status = torch.zeros(B, dtype=torch.int32, device=self.dev)
            _lib.check(L.smpc_people_to_status_device(h, B, self.A, d_raw.data_ptr(), d_np.data_ptr(), init.data_ptr(),
                                                      has_people.data_ptr(), st))
            fa = abi.SmpcFormatArgs()
            fa.n_problems, fa.n_poses, fa.n_blocks = B, n, nb
            fa.n_prev_poses = 0 if self.prev_poses is None else self.prev_poses.shape[1]
            fa.n_prev_cmds = 0 if self.prev_cmds is None else self.prev_cmds.shape[1]
            fa.time_step, fa.current_path_w, fa.current_cmds_w = p.time_step, p.current_path_w, p.current_cmds_w
            fa.poses, fa.cmds, fa.speed = d_poses.data_ptr(), d_cmds.data_ptr(), d_speed.data_ptr()
            fa.prev_poses = None if self.prev_poses is None else self.prev_poses.data_ptr()
            fa.prev_cmds = None if self.prev_cmds is None else self.prev_cmds.data_ptr()
            fa.robot, fa.pose0, fa.u0 = robot.data_ptr(), pose0.data_ptr(), u0.data_ptr()
            fa.path_xy, fa.goal_yaw = path_xy.data_ptr(), goal_yaw.data_ptr()
            _lib.check(L.smpc_format_batch_device(h, C.byref(fa), st))
            pa = abi.SmpcProjectArgs()
            pa.n_problems, pa.n_steps, pa.n_agents, pa.n_grids = B, S, self.A, d_oorg.shape[0]
            pa.od_width, pa.od_height, pa.od_resolution = int(od["width"]), int(od["height"]), float(od["resolution"])
            pa.max_time, pa.time_step = p.max_time, p.time_step
            pa.od_origin, pa.od_indexes = d_oorg.data_ptr(), d_oidx.data_ptr()
            pa.od_index = None if d_osel is None else d_osel.data_ptr()
            pa.robot, pa.people_init, pa.agents, pa.status = (robot.data_ptr(), init.data_ptr(), agents.data_ptr(),
                                                              status.data_ptr())
            _lib.check(L.smpc_project_people_batch_device(h, C.byref(pa), st))
            arrays = dict(pose0=pose0, u0=u0, path_xy=path_xy, goal_yaw=goal_yaw, agents=agents, has_people=has_people,
                          costmaps=d_maps, costmap_origin=d_morg, costmap_index=d_midx)
            bs = abi.make_batch_struct(arrays, B, S, self.A, d_maps.shape[0], d_maps.shape[2], d_maps.shape[1],
                                       float(costmap_resolution), float(np.float32(p.time_step)))
            out = dict(cmds=torch.empty(B, n, 2, dtype=f64, device=self.dev),
                       path=torch.empty(B, n, 3, dtype=f64, device=self.dev),
                       usable=torch.empty(B, dtype=torch.uint8, device=self.dev),
                       termination=torch.empty(B, dtype=torch.int32, device=self.dev),
                       iterations=torch.empty(B, dtype=torch.int32, device=self.dev),
                       cost_final=torch.empty(B, dtype=f64, device=self.dev))
            self.opt.solve_batch_device(bs, out, stream=st)
            if self.prev_poses is None or self.prev_poses.shape[1] != n:
                # first tick: memory = current seed (src/optimizer.cpp:177-181); it is then overwritten where usable
                self.prev_poses = d_poses.clone()
                self.prev_cmds = torch.cat([d_cmds, d_cmds[:, -1:]], dim=1).contiguous()
            _lib.check(L.smpc_memory_update_device(h, B, n, out["usable"].data_ptr(), out["path"].data_ptr(),
                                                   out["cmds"].data_ptr(), self.prev_poses.data_ptr(),
                                                   self.prev_cmds.data_ptr(), st))
        self.stream.synchronize()
        res = {k: v.cpu().numpy() for k, v in out.items()}
        res["optimized"] = res.pop("usable").astype(bool)
        res["people_proj"] = agents.cpu().numpy()
        res["project_status"] = status.cpu().numpy()
        return res
