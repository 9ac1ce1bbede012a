"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the policy forward on the hot path.

* ``policy_mean``  -- ``GaussianNeuralNetwork.compute`` (envs/navigation/learning/skrl/models.py:89-102) with the
  ``HeightmapEncoder`` (:24-36) as built by ``gaussian_model_skrl`` (configure_models.py:37-53):
  ``x = s[:, 0:4]``; ``e = enc(s[:, 3:-1])`` (heading column included, last ray dropped -- quirk kept);
  ``cat`` -> 64 -> 256 -> 160 -> 128 -> 2, LeakyReLU(0.01) between, Tanh at the end.
* ``gaussian_act`` -- skrl 1.1.0 ``GaussianMixin.act`` (third-party, absent from /root/reference; restated from
  SURVEY.md Appendix A.4; ctor args models.py:65-67): clamp log_std to [-20, 2], ``a = mean + exp(log_std) * eps``,
  clamp to the action box [-1, 1] (train.py:134), ``log_prob = sum_j log N(a_j)``.

Pinned by ``tests/golden/policy.npz``: weights of ``best_agent.pt`` and the outputs of the imported reference
network on seeded observations.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

WEIGHT_KEYS = ("dense_encoder.encoder_layers.0", "dense_encoder.encoder_layers.2", "mlp.0", "mlp.2", "mlp.4", "mlp.6")


def policy_mean(obs: torch.Tensor, sd: dict) -> torch.Tensor:
    lin = lambda k, x: F.linear(x, sd[k + ".weight"], sd[k + ".bias"])  # noqa: E731
    act = lambda x: F.leaky_relu(x, 0.01)  # noqa: E731
    x = obs[:, 0:4]
    e = act(lin(WEIGHT_KEYS[1], act(lin(WEIGHT_KEYS[0], obs[:, 3:-1]))))
    h = torch.cat([x, e], dim=1)
    h = act(lin(WEIGHT_KEYS[2], h))
    h = act(lin(WEIGHT_KEYS[3], h))
    h = act(lin(WEIGHT_KEYS[4], h))
    return torch.tanh(lin(WEIGHT_KEYS[5], h))


def value_forward(obs: torch.Tensor, sd: dict) -> torch.Tensor:
    """``DeterministicNeuralNetwork.compute`` (models.py:151-162): the policy's encoder + MLP structure with the
    value weights, final ``nn.Linear(128, 1)`` without activation.  Pinned by ``tests/golden/value.npz``."""
    lin = lambda k, x: F.linear(x, sd[k + ".weight"], sd[k + ".bias"])  # noqa: E731
    act = lambda x: F.leaky_relu(x, 0.01)  # noqa: E731
    e = act(lin(WEIGHT_KEYS[1], act(lin(WEIGHT_KEYS[0], obs[:, 3:-1]))))
    h = torch.cat([obs[:, 0:4], e], dim=1)
    for k in WEIGHT_KEYS[2:5]:
        h = act(lin(k, h))
    return lin(WEIGHT_KEYS[5], h)


def gaussian_act(mean: torch.Tensor, log_std_parameter: torch.Tensor, eps: torch.Tensor):
    log_std = torch.clamp(log_std_parameter, -20.0, 2.0)
    std = log_std.exp()
    actions = mean + std * eps  # Normal(mean, std).rsample() with the standard-normal draw made explicit
    actions = torch.clamp(actions, min=-1.0, max=1.0)
    var = std * std
    log_prob = -((actions - mean) ** 2) / (2 * var) - log_std - math.log(math.sqrt(2 * math.pi))
    return actions, log_prob.sum(dim=-1, keepdim=True)


def load_golden_weights(npz) -> dict:
    return {k.replace("__", "."): torch.from_numpy(npz[k]) for k in npz.files
            if not (k.startswith("in_") or k.startswith("ref_"))}
