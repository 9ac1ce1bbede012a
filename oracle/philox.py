"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based variates of the reset path.

The reference draws its reset variates from torch's global generator (rover_envs/envs/navigation/mdp/
randomizations.py:22, 30; utils/terrains/terrain_importer.py:94-95, 169); a torch generator stream cannot be reproduced
inside a kernel ("RNG parity"), so the CUDA path defines the variates as pure functions of
(seed, step, env, role) -- ``isaac_rover_orbit_b200/csrc/rng.cuh`` -- and this module restates those functions
independently (vectorised numpy, written from the published Philox4x32-10 algorithm of Salmon et al., "Parallel Random
Numbers: As Easy as 1, 2, 3", SC'11 / Random123), pinned by Random123's known-answer vectors
(``tests/test_rng_cpu.py``).  The oracle step consumes the arrays produced here; the kernel generates its own.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
    """``counter [..., 4]`` uint32, ``key = (k0, k1)`` -> ``[..., 4]`` uint32 (10 rounds)."""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def u01(x: np.ndarray) -> np.ndarray:
    """24 random bits -> [0, 1) on the grid of multiples of 2^-24 (exact in fp32)."""
    return ((x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def _key(seed: int):
    return seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF


def _env_stream(seed: int, step: int, n_envs: int, stream: int) -> np.ndarray:
    ctr = np.zeros((n_envs, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(n_envs, dtype=np.uint32)
    ctr[:, 1] = step & 0xFFFFFFFF
    ctr[:, 2] = (step >> 32) & 0xFFFFFFFF
    ctr[:, 3] = stream
    return philox4x32_10(ctr, _key(seed))


def spawn_perm(seed: int, step: int, n_spawns: int, count: int) -> np.ndarray:
    """Row of the spawn table for envs ``0 .. count-1`` (taken if the env resets): cycle-walking hash permutation of
    ``[0, n_spawns)`` evaluated at the env id."""
    ctr = np.array([[0xFFFFFFFF, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, 0xFFFFFFFF]], dtype=np.uint32)
    k = [int(v) for v in philox4x32_10(ctr, _key(seed))[0]]
    m = max(n_spawns - 1, 1)
    for s in (1, 2, 4, 8, 16):
        m |= m >> s
    out = np.empty(count, dtype=np.int64)
    for j in range(count):
        x = j
        while True:
            for r in range(4):
                x = (x + k[r]) & m
                x = (x * 0x9E3779B1) & m
                x ^= x >> 5
                x = (x * 0x85EBCA6B) & m
                x ^= x >> 3
            if x < n_spawns:
                break
        out[j] = x
    return out


def variates(seed: int, step: int, n_envs: int, n_rounds: int, n_spawns: int):
    """``(spawn_by_env [min(N, n_spawns)] int64, yaw_u [N], heading_u [N], theta_u [N, n_rounds])`` of one (seed, step)."""
    s0 = _env_stream(seed, step, n_envs, 0)
    theta = np.empty((n_envs, n_rounds), dtype=np.float32)
    for r0 in range(0, n_rounds, 4):
        w = u01(_env_stream(seed, step, n_envs, 1 + r0 // 4))
        theta[:, r0:r0 + 4] = w[:, : min(4, n_rounds - r0)]
    return spawn_perm(seed, step, n_spawns, min(n_envs, n_spawns)), u01(s0[:, 0]), u01(s0[:, 1]), theta
