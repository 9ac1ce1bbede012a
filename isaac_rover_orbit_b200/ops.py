"""Thin Python operators over the C ABI (``include/rover_b200.h``): tensors in, tensors out, current stream.

No operator has a CPU or eager-PyTorch fallback: CUDA tensors are required and a missing library raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .config import RoverEnvCfg
from .plane_cells import PlaneCells, build_plane_cells
from .scan_grid import ScanGrid, build_scan_grid


# ------------------------------------------------------------------------------------------------------
# height scan
# ------------------------------------------------------------------------------------------------------
class ScanGridHandle:
    """Device-resident home grid + the host struct the launcher reads (``RoverScanGrid``)."""

    def __init__(self, grid: ScanGrid, device, cells: PlaneCells | None = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScanGridHandle needs a CUDA device; there is no CPU fallback")
        self.grid = grid
        self.cells = cells
        self.cells_struct = None
        if cells is not None:
            self.cells_xs = cells.xs.to(self.device).contiguous()
            self.cells_ys = cells.ys.to(self.device).contiguous()
            self.cells_entries = cells.entries.to(self.device).contiguous()
            # the same table as two planes [2, ny, nx, 4] (variant 5 stages windows from it: conflict-free LDS.128)
            self.cells_planar = self.cells_entries.view(cells.ny, cells.nx, 2, 4).permute(2, 0, 1, 3).contiguous()
            self.cells_struct = _lib.PlaneCells(self.cells_xs.data_ptr(), self.cells_ys.data_ptr(),
                                                self.cells_entries.data_ptr(), cells.nx, cells.ny, cells.inv_dx,
                                                cells.inv_dy, self.cells_planar.data_ptr())
        self.cell_start = grid.cell_start.to(self.device).contiguous()
        self.records = grid.records.to(self.device).contiguous()
        if grid.n_records == 0:
            self.records = torch.zeros(1, 12, device=self.device)
        s = _lib.ScanGrid()
        s.n_levels = len(grid.levels)
        s.span = grid.span
        for i, lv in enumerate(grid.levels):
            s.level[i] = _lib.ScanLevel(lv.ox, lv.oy, lv.cell, lv.inv_cell, lv.ncx, lv.ncy, lv.start_offset, 0)
        s.cell_start = self.cell_start.data_ptr()
        s.records = self.records.data_ptr()
        s.n_records = grid.n_records
        self.struct = s

    @classmethod
    def from_mesh(cls, vertices, faces, device, cell_size=None, plane_cells: bool = True) -> "ScanGridHandle":
        v = vertices.detach().cpu().numpy() if isinstance(vertices, torch.Tensor) else np.asarray(vertices)
        f = faces.detach().cpu().numpy() if isinstance(faces, torch.Tensor) else np.asarray(faces)
        grid = build_scan_grid(v, f, cell_size=cell_size)
        cells = build_plane_cells(v, f, fallback_cell=grid.levels[0].cell) if plane_cells else None
        return cls(grid, device, cells)

    def nbytes(self) -> int:
        return self.grid.nbytes() + (self.cells.nbytes() if self.cells is not None else 0)


def grid_pattern(resolution: float = 0.1, size=(3.0, 3.0), offset_pos=(0.0, 0.0, 10.0)) -> torch.Tensor:
    """Local ray starts ``[R,3]`` = ORBIT ``grid_pattern`` + ``RayCasterCfg.offset.pos``.
    x varies fastest: ray ``r = iy * nx + ix``."""
    x = torch.arange(start=-size[0] / 2, end=size[0] / 2 + 1.0e-9, step=resolution)
    y = torch.arange(start=-size[1] / 2, end=size[1] / 2 + 1.0e-9, step=resolution)
    gx, gy = torch.meshgrid(x, y, indexing="xy")
    starts = torch.zeros(gx.numel(), 3, dtype=torch.float32)
    starts[:, 0] = gx.flatten()
    starts[:, 1] = gy.flatten()
    return starts + torch.tensor(offset_pos, dtype=torch.float32)


class RayPattern:
    """ORBIT ``RayCaster.ray_starts`` on the device + its host-side bounding box (the staged kernel's window)."""

    def __init__(self, ray_starts_local: torch.Tensor, device):
        cpu = ray_starts_local.detach().to("cpu", torch.float32).contiguous()
        self.n_rays = cpu.shape[0]
        self.device = torch.device(device)
        self.starts = cpu.to(self.device)
        if self.n_rays:
            self.box = (C.c_float * 4)(float(cpu[:, 0].min()), float(cpu[:, 0].max()), float(cpu[:, 1].min()),
                                       float(cpu[:, 1].max()))
        else:
            self.box = (C.c_float * 4)(0.0, 0.0, 0.0, 0.0)

    @classmethod
    def grid(cls, device, resolution: float = 0.1, size=(3.0, 3.0), offset_pos=(0.0, 0.0, 10.0)) -> "RayPattern":
        return cls(grid_pattern(resolution, size, offset_pos), device)


DEFAULT_SCAN_VARIANT = 5


def height_scan(pos_w: torch.Tensor, quat_w: torch.Tensor, rays, grid: ScanGridHandle,
                max_distance: float = 100.0, base_offset: float = 0.26878, out: torch.Tensor | None = None,
                return_hits: bool = False, variant: int | None = None):
    """``height_scan_rover`` over the CUDA raycaster.  ``out`` may be a ``[N, >=R]`` row-strided view (e.g. the
    scan columns of the observation buffer).  Returns heights ``[N,R]`` (and ``ray_hits_w [N,R,3]``)."""
    if not isinstance(rays, RayPattern):
        rays = RayPattern(rays, pos_w.device)
    ray_starts_local = rays.starts
    variant = DEFAULT_SCAN_VARIANT if variant is None else variant
    dev = _lib.require_cuda(pos_w, quat_w, ray_starts_local)
    n, r = pos_w.shape[0], ray_starts_local.shape[0]
    if pos_w.dtype != torch.float32 or quat_w.dtype != torch.float32 or ray_starts_local.dtype != torch.float32:
        raise RuntimeError("height_scan: fp32 tensors required")
    if pos_w.shape != (n, 3) or quat_w.shape != (n, 4) or ray_starts_local.shape != (r, 3):
        raise RuntimeError("height_scan: bad shapes")
    if grid.device != dev:
        raise RuntimeError("height_scan: grid lives on another device")
    if variant in (2, 3, 4, 5) and grid.cells_struct is None:
        raise RuntimeError("height_scan: variants 2..5 need a ScanGridHandle built with plane_cells=True")
    if out is None:
        out = torch.empty(n, r, dtype=torch.float32, device=dev)
    elif out.dtype != torch.float32 or out.shape != (n, r) or out.stride(1) != 1 or out.device != dev:
        raise RuntimeError("height_scan: out must be fp32 [N,R] with unit inner stride")
    hits = torch.empty(n, r, 3, dtype=torch.float32, device=dev) if return_hits else None
    _lib.check(_lib.load().rover_height_scan(
        _lib.ptr(pos_w), _lib.ptr(quat_w), n, _lib.ptr(ray_starts_local), r, C.byref(rays.box), C.byref(grid.struct),
        C.byref(grid.cells_struct) if grid.cells_struct is not None else None,
        float(max_distance), float(base_offset), C.c_void_p(out.data_ptr()), int(out.stride(0)) if n > 0 else r,
        _lib.ptr(hits), int(variant), _lib.current_stream(dev)))
    return (out, hits) if return_hits else out


def height_scan_obs(pos_w: torch.Tensor, quat_w: torch.Tensor, rays, grid: ScanGridHandle, obs: torch.Tensor,
                    obs_bf16: torch.Tensor, head_cols: int = 4, max_distance: float = 100.0,
                    base_offset: float = 0.26878) -> None:
    """Height scan into the observation buffers of the closed loop: heights -> ``obs[:, head_cols:head_cols + R]``
    (fp32) and the bf16 mirror of ``obs[:, :head_cols + R]`` -> ``obs_bf16`` (``policy.alloc_obs_bf16``), the operand
    of ``GaussianNeuralNetwork.compute_bf16``.  One launch (variant 5 with the extra stores)."""
    if not isinstance(rays, RayPattern):
        rays = RayPattern(rays, pos_w.device)
    _lib.require_cuda(pos_w, quat_w)  # obs / obs_bf16 are row-strided views: checked below
    n = pos_w.shape[0]
    if (not obs.is_cuda or obs.dtype != torch.float32 or obs.shape[0] != n or obs.stride(1) != 1 or obs.shape[1] < head_cols + rays.n_rays
            or obs_bf16.dtype != torch.bfloat16 or obs_bf16.shape[0] != n or obs_bf16.stride(1) != 1
            or obs_bf16.shape[1] < head_cols + rays.n_rays or not obs_bf16.is_cuda):
        raise RuntimeError("height_scan_obs: obs must be fp32 and obs_bf16 bf16, both [N, >= head_cols + R] with unit inner stride")
    dev = pos_w.device
    _lib.check(_lib.load().rover_height_scan_obs(
        _lib.ptr(pos_w), _lib.ptr(quat_w), n, _lib.ptr(rays.starts), rays.n_rays, C.byref(rays.box),
        C.byref(grid.struct), C.byref(grid.cells_struct) if grid.cells_struct is not None else None,
        float(max_distance), float(base_offset), C.c_void_p(obs.data_ptr()), int(obs.stride(0)), int(head_cols),
        C.c_void_p(obs_bf16.data_ptr()), int(obs_bf16.stride(0)), _lib.current_stream(dev)))


# ------------------------------------------------------------------------------------------------------
# fused MDP step
# ------------------------------------------------------------------------------------------------------
def _require_f32(what: str, **tensors) -> None:
    """The kernels read raw fp32: a tensor of another dtype would be reinterpreted silently."""
    for name, t in tensors.items():
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError(f"{what}: {name} must be fp32, got {t.dtype}")


def mdp_params(cfg: RoverEnvCfg) -> _lib.MdpParams:
    a = cfg.actions
    p = _lib.MdpParams()
    p.scale_lin, p.scale_ang = a.scale
    p.offset_lin, p.offset_ang = a.offsets()
    p.wheelbase_length = a.wheelbase_length
    p.middle_wheel_distance = a.middle_wheel_distance
    p.rear_and_front_wheel_distance = a.rear_and_front_wheel_distance
    p.wheel_radius = a.wheel_radius
    p.min_radius = a.middle_wheel_distance * 0.8  # evaluated in double, then fp32 (ackermann_actions.py:264)
    p.wheel_diameter = a.wheel_radius * 2  # ackermann_actions.py:503
    p.action_variant = getattr(a, "variant", 2)
    for i, w in enumerate(cfg.rewards.weights):
        p.weight[i] = w
    p.reached_threshold = cfg.rewards.reached_threshold
    p.far_threshold = cfg.rewards.far_threshold
    p.step_dt = cfg.step_dt
    p.max_episode_length = cfg.max_episode_length
    p.obs_distance_scale = cfg.obs_distance_scale
    p.obs_heading_scale = cfg.obs_heading_scale
    p.target_distance = cfg.commands.target_distance
    p.resampling_time = cfg.commands.resampling_time_range[0]
    p.heading_lo, p.heading_hi = cfg.commands.heading_range
    p.spawn_z_offset = cfg.spawn_z_offset
    p.num_bodies = cfg.num_contact_bodies
    p.episode_length_s = cfg.episode_length_s
    return p


def ackermann(actions: torch.Tensor, params: _lib.MdpParams, variant: int = 2):
    """Stand-alone Ackermann kinematics of any of the reference's three variants.
    Returns ``(processed [N,2], joint_pos [N,4], joint_vel [N,6])`` in the joint orders of that variant."""
    dev = _lib.require_cuda(actions)
    if actions.dtype != torch.float32 or actions.dim() != 2 or actions.shape[1] != 2:
        raise RuntimeError("ackermann: actions must be fp32 [N,2]")
    n = actions.shape[0]
    p = _lib.MdpParams.from_buffer_copy(params)
    p.action_variant = int(variant)
    processed = torch.empty(n, 2, device=dev)
    jp, jv = torch.empty(n, 4, device=dev), torch.empty(n, 6, device=dev)
    _lib.check(_lib.load().rover_ackermann(_lib.ptr(actions), n, C.byref(p), _lib.ptr(processed), _lib.ptr(jp),
                                            _lib.ptr(jv), _lib.current_stream(dev)))
    return processed, jp, jv


@dataclass
class MdpBuffers:
    """Persistent manager state + per-step outputs of the fused kernels, all on one CUDA device."""

    n: int
    device: torch.device
    # state (names follow the reference objects they belong to)
    action: torch.Tensor
    prev_action: torch.Tensor
    pos_cmd_w: torch.Tensor
    heading_cmd_w: torch.Tensor
    pos_cmd_b: torch.Tensor
    heading_cmd_b: torch.Tensor
    time_left: torch.Tensor
    command_counter: torch.Tensor
    episode_length_buf: torch.Tensor
    episode_sums: torch.Tensor
    env_origins: torch.Tensor
    err_pos: torch.Tensor
    err_heading: torch.Tensor
    # outputs
    processed_actions: torch.Tensor
    joint_pos: torch.Tensor
    joint_vel: torch.Tensor
    reward: torch.Tensor
    term_rewards: torch.Tensor
    term_values: torch.Tensor
    terminated: torch.Tensor
    truncated: torch.Tensor
    term_flags: torch.Tensor
    reset_flags: torch.Tensor
    block_reset_counts: torch.Tensor
    spawn_index: torch.Tensor
    stats: torch.Tensor
    scratch: torch.Tensor

    @staticmethod
    def allocate(n: int, device) -> "MdpBuffers":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("MdpBuffers need a CUDA device; there is no CPU fallback")
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)  # noqa: E731
        i64 = lambda *s: torch.zeros(*s, dtype=torch.int64, device=device)  # noqa: E731
        u8 = lambda *s: torch.zeros(*s, dtype=torch.uint8, device=device)  # noqa: E731
        blocks = (n + _lib.MDP_BLOCK - 1) // _lib.MDP_BLOCK
        return MdpBuffers(
            n, device, f(n, 2), f(n, 2), f(n, 3), f(n), f(n, 3), f(n), f(n), i64(n), i64(n), f(n, 7), f(n, 3), f(n),
            f(n), f(n, 2), f(n, 4), f(n, 6), f(n), f(n, 7), f(n, 7), u8(n), u8(n), u8(n, 4), u8(n),
            torch.zeros(max(blocks, 1), dtype=torch.int32, device=device), torch.full((n,), -1, dtype=torch.int64,
                                                                                      device=device),
            f(_lib.STATS_LEN), f(max(blocks, 1) * _lib.STATS_LEN + 1))

    @property
    def lookback(self) -> torch.Tensor:
        """Descriptors + ticket + epoch of the fused step's look-back (``mdp_step``); allocated on first use."""
        lb = self.__dict__.get("_lookback")
        if lb is None:
            blocks = (self.n + _lib.MDP_BLOCK - 1) // _lib.MDP_BLOCK
            lb = self.__dict__["_lookback"] = torch.zeros(max(blocks, 1) + 2, dtype=torch.int64, device=self.device)
        return lb

    def state_struct(self) -> _lib.MdpState:
        """Host struct of device pointers; built once (the buffers are never re-allocated)."""
        st = self.__dict__.get("_state_struct")
        if st is None:
            st = self.__dict__["_state_struct"] = _lib.MdpState(*[getattr(self, k).data_ptr() for k in _lib._STATE_FIELDS])
        return st

    def out_struct(self) -> _lib.MdpOut:
        out = self.__dict__.get("_out_struct")
        if out is None:
            out = self.__dict__["_out_struct"] = _lib.MdpOut(*[getattr(self, k).data_ptr() for k in _lib._OUT_FIELDS])
        return out


class TerrainTablesHandle:
    """Device tables + host struct (``RoverTerrainTables``) for the resample path."""

    def __init__(self, heightmap, safe_mask, offset_xy, spawn_table, resolution, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TerrainTablesHandle needs a CUDA device; there is no CPU fallback")
        self.heightmap = heightmap.to(self.device, torch.float32).contiguous()
        mask = safe_mask.reshape(safe_mask.shape[0], safe_mask.shape[1])
        self.safe_mask = mask.to(self.device, torch.uint8).contiguous()
        self.spawn_table = spawn_table.to(self.device, torch.float32).contiguous()
        off = [float(v) for v in offset_xy]
        self.struct = _lib.TerrainTables(self.heightmap.data_ptr(), self.safe_mask.data_ptr(),
                                         self.heightmap.shape[0], self.heightmap.shape[1], off[0], off[1],
                                         float(resolution), self.spawn_table.data_ptr(), self.spawn_table.shape[0], 0)
        if self.safe_mask.shape != self.heightmap.shape:
            raise RuntimeError("safe_mask and heightmap shapes differ")


def mdp_pre_step(buf: MdpBuffers, params: _lib.MdpParams, actions: torch.Tensor | None,
                 force_matrix_w: torch.Tensor | None, phases: int = _lib.PRE_ALL):
    """Action shift + Ackermann (``PRE_ACTIONS``) and counters + terminations + rewards (``PRE_TERMS``);
    one launch, in place on ``buf``."""
    dev = _lib.require_cuda(actions, force_matrix_w)
    if dev is not None and dev != buf.device:
        raise RuntimeError("mdp_pre_step: inputs on another device than the buffers")
    if phases & _lib.PRE_ACTIONS and (actions is None or actions.shape != (buf.n, 2) or actions.dtype != torch.float32):
        raise RuntimeError("mdp_pre_step: actions must be fp32 [N,2]")
    if phases & _lib.PRE_TERMS and (force_matrix_w is None or force_matrix_w.dtype != torch.float32 or
                                    force_matrix_w.numel() != buf.n * params.num_bodies * 3):
        raise RuntimeError("mdp_pre_step: force_matrix_w must be fp32 [N,B,1,3]")
    st, out = buf.state_struct(), buf.out_struct()
    _lib.check(_lib.load().rover_mdp_pre_step(_lib.ptr(actions), _lib.ptr(force_matrix_w), buf.n, C.byref(params),
                                               C.byref(st), C.byref(out), int(phases),
                                               _lib.current_stream(buf.device)))


def mdp_post_step(buf: MdpBuffers, params: _lib.MdpParams, tables: TerrainTablesHandle, root_pos_w: torch.Tensor,
                  root_quat_w: torch.Tensor, spawn_perm: torch.Tensor, yaw_u: torch.Tensor, heading_u: torch.Tensor,
                  theta_u: torch.Tensor, obs: torch.Tensor | None = None, phases: int = _lib.PHASE_ALL,
                  xchg=None):
    """Reset / resample / command update / observation head (one launch).  ``root_*`` are updated in place
    for the reset envs; ``buf.stats`` is accumulated; ``buf.spawn_index`` holds the spawn rows used (-1 else)."""
    dev = _lib.require_cuda(root_pos_w, root_quat_w, spawn_perm, yaw_u, heading_u, theta_u)
    if dev != buf.device or tables.device != dev:
        raise RuntimeError("mdp_post_step: tensors on different devices")
    _require_f32("mdp_post_step", root_pos_w=root_pos_w, root_quat_w=root_quat_w, yaw_u=yaw_u, heading_u=heading_u,
                 theta_u=theta_u)
    n = buf.n
    if root_pos_w.shape != (n, 3) or root_quat_w.shape != (n, 4):
        raise RuntimeError("mdp_post_step: bad root state shapes")
    if spawn_perm.dtype != torch.int64 or spawn_perm.numel() < n:
        raise RuntimeError("mdp_post_step: spawn_perm must be int64 with at least N entries")
    if theta_u.dim() != 2 or theta_u.shape[0] != n or yaw_u.shape != (n,) or heading_u.shape != (n,):
        raise RuntimeError("mdp_post_step: variates must be yaw_u[N], heading_u[N], theta_u[N,R]")
    obs_ptr, obs_stride = None, 0
    if obs is not None:
        if obs.dtype != torch.float32 or obs.shape[0] != n or obs.stride(1) != 1 or obs.shape[1] < 4:
            raise RuntimeError("mdp_post_step: obs must be fp32 [N,>=4] with unit inner stride")
        obs_ptr, obs_stride = C.c_void_p(obs.data_ptr()), int(obs.stride(0))
    st, out = buf.state_struct(), buf.out_struct()
    # xchg: a dist.P2PStats -- the launch also publishes the rank's running statistics to every rank's mailbox
    _lib.check(_lib.load().rover_mdp_post_step_x(
        _lib.ptr(root_pos_w), _lib.ptr(root_quat_w), n, C.byref(params), C.byref(st), C.byref(out),
        C.byref(tables.struct), _lib.ptr(spawn_perm), _lib.ptr(yaw_u), _lib.ptr(heading_u), _lib.ptr(theta_u),
        int(theta_u.shape[1]), _lib.ptr(buf.spawn_index), _lib.ptr(buf.stats), _lib.ptr(buf.scratch), obs_ptr,
        obs_stride, int(phases), C.byref(xchg.struct) if xchg is not None else None, _lib.current_stream(dev)))


def mdp_step(buf: MdpBuffers, params: _lib.MdpParams, tables: "TerrainTablesHandle", new_actions: torch.Tensor | None,
             force_matrix_w: torch.Tensor | None, root_pos_w: torch.Tensor, root_quat_w: torch.Tensor,
             spawn_perm: torch.Tensor, yaw_u: torch.Tensor, heading_u: torch.Tensor, theta_u: torch.Tensor,
             obs: torch.Tensor | None = None, pre_phases: int = _lib.PRE_ACTIONS | _lib.PRE_TERMS,
             phases: int = _lib.PHASE_ALL, xchg=None):
    """``mdp_pre_step`` + ``mdp_post_step`` in ONE launch (``rover_mdp_step``): for loops whose physics does not sit
    between the two.  Bit-identical to the two-launch sequence (``tests/test_gpu_parity.py``)."""
    dev = _lib.require_cuda(root_pos_w, root_quat_w, spawn_perm, yaw_u, heading_u, theta_u)
    if dev != buf.device or tables.device != dev:
        raise RuntimeError("mdp_step: tensors on different devices")
    _require_f32("mdp_step", root_pos_w=root_pos_w, root_quat_w=root_quat_w, yaw_u=yaw_u, heading_u=heading_u,
                 theta_u=theta_u, new_actions=new_actions, force_matrix_w=force_matrix_w)
    n = buf.n
    if root_pos_w.shape != (n, 3) or root_quat_w.shape != (n, 4):
        raise RuntimeError("mdp_step: bad root state shapes")
    if (pre_phases & _lib.PRE_ACTIONS) and (new_actions is None or new_actions.shape != (n, 2)):
        raise RuntimeError("mdp_step: new_actions must be [N,2]")
    if (pre_phases & _lib.PRE_TERMS) and (force_matrix_w is None or force_matrix_w.numel() != n * params.num_bodies * 3):
        raise RuntimeError("mdp_step: force_matrix_w must be [N, num_bodies, 1, 3]")
    for t in (new_actions, force_matrix_w):
        if t is not None:
            _lib.require_cuda(t)
    if spawn_perm.dtype != torch.int64 or spawn_perm.numel() < n:
        raise RuntimeError("mdp_step: spawn_perm must be int64 with at least N entries")
    if theta_u.dim() != 2 or theta_u.shape[0] != n or yaw_u.shape != (n,) or heading_u.shape != (n,):
        raise RuntimeError("mdp_step: variates must be yaw_u[N], heading_u[N], theta_u[N,R]")
    obs_ptr, obs_stride = None, 0
    if obs is not None:
        if obs.dtype != torch.float32 or obs.shape[0] != n or obs.stride(1) != 1 or obs.shape[1] < 4:
            raise RuntimeError("mdp_step: obs must be fp32 [N,>=4] with unit inner stride")
        obs_ptr, obs_stride = C.c_void_p(obs.data_ptr()), int(obs.stride(0))
    st, out = buf.state_struct(), buf.out_struct()
    _lib.check(_lib.load().rover_mdp_step(
        _lib.ptr(new_actions), _lib.ptr(force_matrix_w), _lib.ptr(root_pos_w), _lib.ptr(root_quat_w), n, C.byref(params),
        C.byref(st), C.byref(out), C.byref(tables.struct), _lib.ptr(spawn_perm), _lib.ptr(yaw_u), _lib.ptr(heading_u),
        _lib.ptr(theta_u), int(theta_u.shape[1]), _lib.ptr(buf.spawn_index), _lib.ptr(buf.stats), _lib.ptr(buf.scratch),
        _lib.ptr(buf.lookback), obs_ptr, obs_stride, int(pre_phases), int(phases),
        C.byref(xchg.struct) if xchg is not None else None, _lib.current_stream(dev)))
