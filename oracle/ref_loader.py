"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference term functions.

Imports the reference's own first-party torch code *by path* from ``/root/reference``
under a stub tree for the third-party packages it needs but that are not installed
(ORBIT ``omni.isaac.orbit.*``, ``carb``, ``skrl``, ``pymeshlab``, ``matplotlib``, USD helpers).
Nothing is copied: the reference sources stay where they are and are executed in place.

Used only by ``tests/golden/make_golden.py`` (fixture generation, run in the build container)
and by CPU tests that cross-check the restated oracle when ``/root/reference`` is present.
It must never be imported by the product package, ``bench.py``'s GPU arm or ``smoke()``:
``/root/reference`` does not exist on the GPU box.

The third-party semantics the stubs need (ORBIT math helpers, ActionTerm/CommandTerm base
behaviour) are restated in :mod:`oracle.orbit_math` / :mod:`oracle.managers`
(SURVEY.md Appendix A); the stubs below only forward to those restatements.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ROVER_REFERENCE_ROOT", "/root/reference")

_LOADED: dict[str, types.ModuleType] = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rover_envs"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package so that sub-imports resolve
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(_mod(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs() -> None:
    """Install the stub tree into ``sys.modules`` (idempotent)."""
    if "omni.isaac.orbit" in sys.modules and getattr(sys.modules["omni.isaac.orbit"], "_rover_stub", False):
        return
    import torch

    from . import orbit_math

    # ---- ORBIT base classes (behaviour restated from SURVEY.md Appendix A.2) ----
    class ActionTerm:
        """Stub of ORBIT ``ActionTerm``: stores cfg/env, resolves the asset (A.2)."""

        def __init__(self, cfg, env):
            self.cfg = cfg
            self._env = env
            self._asset = env.scene[cfg.asset_name]

        @property
        def num_envs(self):
            return self._env.num_envs

        @property
        def device(self):
            return self._env.device

    class ActionTermCfg:
        class_type = None
        asset_name = None

    class CommandTerm:
        """Stub of ORBIT ``CommandTerm`` (A.2): metrics dict, time_left, counter, compute/reset."""

        def __init__(self, cfg, env):
            self.cfg = cfg
            self._env = env
            self.metrics = {}
            self.time_left = torch.zeros(env.num_envs, device=env.device)
            self.command_counter = torch.zeros(env.num_envs, dtype=torch.long, device=env.device)

        @property
        def num_envs(self):
            return self._env.num_envs

        @property
        def device(self):
            return self._env.device

    class CommandTermCfg:
        class_type = None
        resampling_time_range = None
        debug_vis = False

    class TerrainImporter:
        def __init__(self, cfg):
            self.cfg = cfg
            self.device = getattr(cfg, "device", "cpu")
            self.env_origins = torch.zeros(cfg.num_envs, 3, device=self.device)

    class TerrainImporterCfg:
        pass

    class _Anything:
        def __init__(self, *a, **k):
            pass

    def configclass(cls):
        return cls

    class SceneEntityCfg:
        def __init__(self, name=None, **kw):
            self.name = name
            self.__dict__.update(kw)

    omni = _mod("omni")
    _mod("omni.isaac")
    orbit = _mod("omni.isaac.orbit", _rover_stub=True)
    _mod("omni.isaac.orbit.assets", Articulation=_Anything, RigidObject=_Anything, ArticulationCfg=_Anything)
    _mod("omni.isaac.orbit.assets.articulation", Articulation=_Anything)
    _mod("omni.isaac.orbit.envs", BaseEnv=_Anything, RLTaskEnv=_Anything)
    _mod("omni.isaac.orbit.envs.mdp")
    _mod("omni.isaac.orbit.envs.mdp.commands")
    _mod("omni.isaac.orbit.envs.mdp.commands.commands_cfg", CommandTermCfg=CommandTermCfg)
    _mod("omni.isaac.orbit.managers", SceneEntityCfg=SceneEntityCfg, CommandTerm=CommandTerm,
         ActionTerm=ActionTerm, ActionTermCfg=ActionTermCfg)
    _mod("omni.isaac.orbit.managers.action_manager", ActionTerm=ActionTerm, ActionTermCfg=ActionTermCfg)
    _mod("omni.isaac.orbit.sensors", RayCaster=_Anything, ContactSensor=_Anything)
    _mod("omni.isaac.orbit.utils", configclass=configclass)
    _mod("omni.isaac.orbit.utils.math", quat_rotate_inverse=orbit_math.quat_rotate_inverse,
         wrap_to_pi=orbit_math.wrap_to_pi, yaw_quat=orbit_math.yaw_quat)
    _mod("omni.isaac.orbit.markers", VisualizationMarkers=_Anything)
    _mod("omni.isaac.orbit.markers.config", CUBOID_MARKER_CFG=_Anything())
    _mod("omni.isaac.orbit.terrains", TerrainImporter=TerrainImporter, TerrainImporterCfg=TerrainImporterCfg)
    del omni, orbit

    _mod("carb", log_info=lambda *a, **k: None, log_error=lambda *a, **k: None)

    # ---- skrl 1.1.0 mixins (A.4): only what GaussianNeuralNetwork.__init__/compute touches ----
    class Model(torch.nn.Module):
        def __init__(self, observation_space, action_space, device=None):
            super().__init__()
            self.observation_space = observation_space
            self.action_space = action_space
            self.device = device

    class GaussianMixin:
        def __init__(self, clip_actions=False, clip_log_std=True, min_log_std=-20.0, max_log_std=2.0,
                     reduction="sum", role=""):
            self._clip_actions = clip_actions
            self._clip_log_std = clip_log_std
            self._log_std_min = min_log_std
            self._log_std_max = max_log_std
            self._reduction = reduction

    class DeterministicMixin:
        def __init__(self, clip_actions=False, role=""):
            self._clip_actions = clip_actions

    _mod("skrl")
    _mod("skrl.models")
    _mod("skrl.models.torch")
    _mod("skrl.models.torch.base", Model=Model)
    _mod("skrl.models.torch.gaussian", GaussianMixin=GaussianMixin)
    _mod("skrl.models.torch.deterministic", DeterministicMixin=DeterministicMixin)

    # ---- skrl 1.1.0 trainer base (third-party, absent here): restated from the published 1.1.0 sources --
    # skrl/trainers/torch/base.py (Trainer.__init__, single_agent_eval) and sequential.py (default config)
    class Agent:
        pass

    class Wrapper:
        pass

    class Trainer:
        def __init__(self, env, agents, agents_scope=None, cfg=None):
            self.cfg = cfg if cfg is not None else {}
            self.env = env
            self.agents = agents
            self.agents_scope = agents_scope if agents_scope is not None else []
            self.timesteps = self.cfg.get("timesteps", 0)
            self.headless = self.cfg.get("headless", False)
            self.disable_progressbar = self.cfg.get("disable_progressbar", False)
            self.close_environment_at_exit = self.cfg.get("close_environment_at_exit", True)
            self.initial_timestep = 0
            self.num_simultaneous_agents = len(agents) if isinstance(agents, (list, tuple)) else 1

        def single_agent_eval(self):
            states, infos = self.env.reset()
            for timestep in range(self.initial_timestep, self.timesteps):
                with torch.no_grad():
                    actions = self.agents.act(states, timestep=timestep, timesteps=self.timesteps)[0]
                next_states, rewards, terminated, truncated, infos = self.env.step(actions)
                if not self.headless:
                    self.env.render()
                with torch.no_grad():
                    super(type(self.agents), self.agents).record_transition(
                        states=states, actions=actions, rewards=rewards, next_states=next_states,
                        terminated=terminated, truncated=truncated, infos=infos, timestep=timestep,
                        timesteps=self.timesteps)
                if self.env.num_envs > 1:
                    states = next_states
                else:
                    if terminated.any() or truncated.any():
                        with torch.no_grad():
                            states, infos = self.env.reset()
                    else:
                        states = next_states

    _mod("skrl.agents")
    _mod("skrl.agents.torch", Agent=Agent)
    _mod("skrl.envs")
    _mod("skrl.envs.wrappers")
    _mod("skrl.envs.wrappers.torch", Wrapper=Wrapper, wrap_env=lambda env, wrapper=None: env)
    _mod("skrl.trainers")
    _mod("skrl.trainers.torch", Trainer=Trainer)
    _mod("skrl.trainers.torch.sequential",
         SEQUENTIAL_TRAINER_DEFAULT_CONFIG={"timesteps": 100000, "headless": False, "disable_progressbar": False,
                                            "close_environment_at_exit": True})

    if "pymeshlab" not in sys.modules:
        _mod("pymeshlab")
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        _mod("matplotlib")
        _mod("matplotlib.pyplot")

    # first-party helper that needs USD; the path never calls it in this harness
    _mod("rover_envs")
    _mod("rover_envs.envs")
    _mod("rover_envs.envs.navigation")
    _mod("rover_envs.envs.navigation.utils")
    _mod("rover_envs.envs.navigation.utils.terrains")
    _mod("rover_envs.envs.navigation.utils.terrains.usd_utils",
         get_triangles_and_vertices_from_prim=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no USD")))


_FILES = {
    "ackermann_actions": ("rover_envs/mdp/actions/ackermann_actions.py", "rover_envs.mdp.actions.ackermann_actions"),
    "observations": ("rover_envs/envs/navigation/mdp/observations.py", "rover_envs.envs.navigation.mdp.observations"),
    "rewards": ("rover_envs/envs/navigation/mdp/rewards.py", "rover_envs.envs.navigation.mdp.rewards"),
    "terminations": ("rover_envs/envs/navigation/mdp/terminations.py", "rover_envs.envs.navigation.mdp.terminations"),
    "terrain_utils": ("rover_envs/envs/navigation/utils/terrains/terrain_utils.py",
                      "rover_envs.envs.navigation.utils.terrains.terrain_utils"),
    "terrain_importer": ("rover_envs/envs/navigation/utils/terrains/terrain_importer.py",
                         "rover_envs.envs.navigation.utils.terrains.terrain_importer"),
    "randomizations": ("rover_envs/envs/navigation/mdp/randomizations.py",
                       "rover_envs.envs.navigation.mdp.randomizations"),
    "models": ("rover_envs/envs/navigation/learning/skrl/models.py", "rover_envs.envs.navigation.learning.skrl.models"),
    "skrl_utils": ("rover_envs/utils/skrl_utils.py", "rover_envs.utils.skrl_utils"),
}


def load(name: str) -> types.ModuleType:
    """Execute one reference file in place (by path) and return the module."""
    if name in _LOADED:
        return _LOADED[name]
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    install_stubs()
    rel, modname = _FILES[name]
    if name in ("terrain_importer", "randomizations"):
        load("terrain_utils")
    if name == "randomizations":
        load("terrain_importer")
    path = os.path.join(REFERENCE_ROOT, rel)
    spec = importlib.util.spec_from_file_location(modname, path)
    module = importlib.util.module_from_spec(spec)
    # make relative imports inside the file (``from .terrain_utils import``) resolve
    parent = modname.rpartition(".")[0]
    _mod(parent)
    module.__package__ = parent
    sys.modules[modname] = module
    spec.loader.exec_module(module)
    setattr(sys.modules[parent], modname.rpartition(".")[2], module)
    _LOADED[name] = module
    return module


POLICY_CHECKPOINT = os.path.join(
    REFERENCE_ROOT, "rover_envs/envs/navigation/robots/aau_rover/policies/best_agent.pt")
