"""``AckermannAction2`` -- the action term selected by ``AckermannActionCfg.class_type``
(rover_envs/mdp/actions/actions_cfg.py:17; class at rover_envs/mdp/actions/ackermann_actions.py:162-322).

Same constructor, properties and methods; ``process_actions`` + ``apply_actions`` are served by the ACTIONS phase of
the fused pre-step kernel (action-manager shift, scale/offset, Ackermann kinematics in one launch).  The reference
recomputes the identical kinematics ``decimation`` = 6 times per env step (rover_env.py:64-66); here the first
``apply_actions`` after a ``process_actions`` launches, the rest re-bind the same targets.
"""
from __future__ import annotations

import torch

from .. import _lib, ops
from ..config import AckermannActionCfg  # noqa: F401  (re-exported like actions_cfg.AckermannActionCfg)


class AckermannAction2:
    cfg: AckermannActionCfg
    VARIANT = 2

    def __init__(self, cfg: AckermannActionCfg, env):
        self.cfg = cfg
        self._env = env
        self._asset = env.scene[cfg.asset_name]
        self._drive_joint_ids, self._drive_joint_names = self._asset.find_joints(cfg.drive_joint_names)
        self._steering_joint_ids, self._steering_joint_names = self._asset.find_joints(cfg.steering_joint_names)
        self._raw_actions = torch.zeros(self.num_envs, self.action_dim, device=self.device)
        # the term's own copy of the kernel parameters: its class selects the kinematics, the env's struct is not touched
        self._params = _lib.MdpParams.from_buffer_copy(env._params)
        self._params.action_variant = self.VARIANT

    @property
    def num_envs(self) -> int:
        return self._env.num_envs

    @property
    def device(self):
        return self._env.device

    @property
    def action_dim(self) -> int:
        return 2  # (linear velocity, angular velocity)

    @property
    def raw_actions(self) -> torch.Tensor:
        return self._raw_actions

    @property
    def processed_actions(self) -> torch.Tensor:
        return self._env._buf.processed_actions

    def stage_actions(self, actions: torch.Tensor) -> None:
        """``_raw_actions[:] = actions`` only (ackermann_actions.py:227): ``RoverEnv.step`` then runs the rest of
        ``process_actions`` inside its single fused launch."""
        if actions is not self._raw_actions:
            self._raw_actions[:] = actions
        self._env._terms_current = False

    def process_actions(self, actions: torch.Tensor, fused_terms_force: torch.Tensor | None = None):
        """ackermann_actions.py:226-229 (+ ORBIT ActionManager.process_action: prev_action <- action <- actions).
        ``fused_terms_force`` (``RoverEnv.step`` when nothing has to run between the action term and the reward /
        termination terms): the contact forces -- the same launch then also advances the counters and computes the
        terminations and rewards."""
        if actions is not self._raw_actions:  # (the captured step of RoverEnv writes its input straight into raw_actions)
            self._raw_actions[:] = actions
        phases = _lib.PRE_ACTIONS | (_lib.PRE_TERMS if fused_terms_force is not None else 0)
        ops.mdp_pre_step(self._env._buf, self._params, self._raw_actions, fused_terms_force, phases=phases)
        self._env._terms_current = False  # the reward / termination columns describe the previous action until TERMS runs

    def apply_actions(self):
        """ackermann_actions.py:231-236: joint targets [FL,RL,RR,FR] (rad) and [ML,FL,RL,RR,MR,FR] (rad/s).  The
        reference recomputes them on each of the ``decimation`` calls; here they were computed once by
        ``process_actions`` and are re-bound."""
        b = self._env._buf
        self._asset.set_joint_velocity_target(b.joint_vel, joint_ids=self._drive_joint_ids)
        self._asset.set_joint_position_target(b.joint_pos, joint_ids=self._steering_joint_ids)


class AckermannAction(AckermannAction2):
    """``AckermannAction`` (ackermann_actions.py:19-158): turning-point model with the AAU wheel locations hard-coded;
    joint targets come back as [FL,FR,RL,RR] / [FL,FR,ML,MR,RL,RR].  Same ActionTerm API as ``AckermannAction2``."""

    VARIANT = 1


class AckermannAction3:
    """``AckermannAction3`` (ackermann_actions.py:329-420) over ``ackermann()`` (:423-505): a stand-alone controller
    (not an ActionTerm) constructed as ``(cfg, robot, num_envs, device)``; joint ids are re-ordered by
    ``cfg.steering_order`` / ``cfg.drive_order`` like the reference does."""

    def __init__(self, cfg: AckermannActionCfg, robot, num_envs: int, device):
        from ..config import RoverEnvCfg

        self.cfg, self.device, self.num_envs, self._asset = cfg, torch.device(device), num_envs, robot
        self._drive_joint_ids, self._drive_joint_names = robot.find_joints(cfg.drive_joint_names)
        self._steering_joint_ids, self._steering_joint_names = robot.find_joints(cfg.steering_joint_names)
        so, do = list(cfg.steering_order), list(cfg.drive_order)
        # the reference's drive_order spells the middle wheels "CL"/"CR" (actions_cfg.py:49); the AAU joints are "ML"/"MR"
        key = lambda order: (lambda name: order.index({"ML": "CL", "MR": "CR"}.get(name[:2], name[:2])))  # noqa: E731
        self._sorted_steering_ids = [i for _, i in sorted(zip(self._steering_joint_names, self._steering_joint_ids),
                                                          key=lambda t: key(so)(t[0]))]
        self._sorted_drive_ids = [i for _, i in sorted(zip(self._drive_joint_names, self._drive_joint_ids),
                                                       key=lambda t: key(do)(t[0]))]
        self._raw_actions = torch.zeros(num_envs, 2, device=self.device)
        self._processed_actions = torch.zeros_like(self._raw_actions)
        self._params = ops.mdp_params(RoverEnvCfg(num_envs=num_envs, actions=cfg))

    @property
    def action_dim(self) -> int:
        return 2

    @property
    def raw_actions(self) -> torch.Tensor:
        return self._raw_actions

    @property
    def processed_actions(self) -> torch.Tensor:
        return self._processed_actions

    def process_actions(self, actions):
        self._raw_actions[:] = actions

    def apply_actions(self):
        self._processed_actions, self._joint_pos, self._joint_vel = ops.ackermann(self._raw_actions, self._params, 3)
        self._asset.set_joint_velocity_target(self._joint_vel, joint_ids=self._sorted_drive_ids)
        self._asset.set_joint_position_target(self._joint_pos, joint_ids=self._sorted_steering_ids)
