"""Fake env / agent that log every call the trainer makes -- shared by tests/golden/make_trainer_golden.py (driving the
unmodified reference trainer) and tests/test_trainer_cpu.py (driving ours)."""
import torch


class BaseAgent:
    """stands for skrl's base Agent: evaluation records through ``super(type(agent), agent)``"""

    def record_transition(self, **kw):
        self.log.append(["base.record_transition", kw["timestep"], kw["timesteps"], sorted(kw)])


class FakeAgent(BaseAgent):
    def __init__(self, log):
        self.log = log

    def init(self, trainer_cfg=None):
        self.log.append(["init", sorted(trainer_cfg)])

    def set_running_mode(self, mode):
        self.log.append(["set_running_mode", mode])

    def pre_interaction(self, timestep, timesteps):
        self.log.append(["pre_interaction", timestep, timesteps])

    def post_interaction(self, timestep, timesteps):
        self.log.append(["post_interaction", timestep, timesteps])

    def act(self, states, timestep, timesteps):
        self.log.append(["act", timestep, timesteps, float(states.sum())])
        return states[:, :2] * 0.5 + timestep, None, {}

    def record_transition(self, **kw):
        self.log.append(["record_transition", kw["timestep"], kw["timesteps"], sorted(kw), float(kw["states"].sum()),
                         float(kw["next_states"].sum()), float(kw["rewards"].sum())])

    def track_data(self, tag, value):
        self.log.append(["track_data", tag, round(float(value), 6)])


class FakeEnv:
    num_agents = 1

    def __init__(self, log, num_envs=3):
        self.log = log
        self.num_envs = num_envs
        self.t = 0
        self._obs = torch.zeros(num_envs, 4)

    def reset(self):
        self.log.append(["reset"])
        self.t = 0
        return torch.ones(self.num_envs, 4), {}

    def step(self, actions):
        self.t += 1
        self.log.append(["step", round(float(actions.sum()), 6)])
        self._obs = torch.full((self.num_envs, 4), float(self.t)) + actions.sum() * 0.01
        infos = {}
        if self.t % 2 == 0:  # the reference logs one-element tensors only
            infos["episode"] = {"Episode Reward/collision": torch.tensor(-0.25 * self.t), "vector": torch.ones(3),
                                "not a tensor": 1.0}
        return (self._obs, torch.full((self.num_envs, 1), 0.1 * self.t), torch.zeros(self.num_envs, 1, dtype=torch.bool),
                torch.zeros(self.num_envs, 1, dtype=torch.bool), infos)

    def render(self):
        self.log.append(["render"])
