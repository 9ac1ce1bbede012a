"""Argument checks shared by the reward / termination / observation mirrors.

The fused pre-step kernel computes every term once per step with the thresholds of ``env.cfg``; the manager-term
functions hand out columns of that result.  A caller that passes another threshold, another command, or asks before the
terms of the current state exist would silently get numbers for a question it did not ask -- so those calls raise."""
from __future__ import annotations

import numpy as np


def require_current(env, what: str) -> None:
    """The term columns must belong to the current state: ``RoverEnv.step`` (or a ``PRE_TERMS`` launch) computed them
    and no action has been processed since."""
    if not getattr(env, "_terms_current", False):
        raise RuntimeError(f"{what}: the fused term kernel has not run for the current state -- call env.step() (or "
                           f"ops.mdp_pre_step with PRE_TERMS) first; the columns held now describe an earlier step")


def require_command(env, command_name: str, what: str):
    """Like ``env.command_manager.get_command(name)`` in the reference: unknown names raise ``KeyError``."""
    try:
        return env.command_manager.get_command(command_name)
    except KeyError:
        raise KeyError(f"{what}: unknown command {command_name!r}") from None


def require_threshold(given: float, compiled: float, what: str, field: str) -> None:
    if np.float32(given) != np.float32(compiled):
        raise ValueError(f"{what}: threshold={given} differs from the value the fused kernel ran with "
                         f"(cfg.rewards.{field}={compiled}); set it in the config before constructing the env")
