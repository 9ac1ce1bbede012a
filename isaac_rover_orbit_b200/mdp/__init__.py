"""Manager-term functions and classes with the reference's names and signatures
(rover_envs/mdp/actions, rover_envs/envs/navigation/mdp, .../utils/terrains/terrain_importer.py).

Every public name is listed explicitly: this namespace is what a term config (``ObsTerm(func=mdp.…)``, ``RewTerm``,
``DoneTerm``, ``RandTerm``, ``class_type=…``) resolves against when the package replaces the reference's ``mdp``."""
from . import actions, commands, observations, randomizations, rewards, terminations
from .actions import AckermannAction, AckermannAction2, AckermannAction3, AckermannActionCfg
from .commands import TerrainBasedPositionCommand
from .observations import angle_to_target_observation, distance_to_target_euclidean, height_scan_rover, last_action
from .randomizations import reset_root_state_rover
from .rewards import (angle_to_target_penalty, collision_penalty, distance_to_target_reward, far_from_target_reward,
                      heading_soft_contraint, oscillation_penalty, reached_target)
from .terminations import collision_with_obstacles, far_from_target, is_success, time_out

__all__ = [
    "AckermannAction", "AckermannAction2", "AckermannAction3", "AckermannActionCfg", "TerrainBasedPositionCommand",
    *observations.__all__, *randomizations.__all__, *rewards.__all__, *terminations.__all__,
]
