"""Manager-term functions and classes with the reference's names and signatures
(rover_envs/mdp/actions, rover_envs/envs/navigation/mdp, .../utils/terrains/terrain_importer.py)."""
from .actions import AckermannAction, AckermannAction2, AckermannAction3, AckermannActionCfg  # noqa: F401
from .commands import TerrainBasedPositionCommand  # noqa: F401
from .observations import *  # noqa: F401,F403
from .randomizations import *  # noqa: F401,F403
from .rewards import *  # noqa: F401,F403
from .terminations import *  # noqa: F401,F403
