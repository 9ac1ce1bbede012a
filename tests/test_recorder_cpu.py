"""CPU: the offline-dataset recorder (isaac_rover_orbit_b200/recorder.py: staging ring, batched transfer, regrouping by done
events) writes, call for call into the h5py API, the same files as the reference's per-rover Python loop
(oracle/recorder.py restates base.py:42-67 + hdf_recorder.py:34-88)."""
import types

import numpy as np
import pytest
import torch

from isaac_rover_orbit_b200.recorder import DataRecorderBase, HDF5DataRecorder, SequentialCollector
from oracle.recorder import FakeH5, reference_recorder_run


def _env(num_obs, num_act):
    return types.SimpleNamespace(observation_space=types.SimpleNamespace(shape=(num_obs,), dtype=np.float32),
                                 action_space=types.SimpleNamespace(shape=(num_act,), dtype=np.float32))


def _steps(n_envs, n_steps, num_obs, num_act, p_done, seed, extras):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n_steps):
        info = {k: rng.standard_normal((n_envs, *v["shape"])).astype(v["dtype"]) for k, v in extras.items()}
        out.append((rng.standard_normal((n_envs, num_obs)).astype(np.float32), rng.standard_normal((n_envs, num_act)).astype(np.float32),
                    rng.standard_normal(n_envs).astype(np.float32), rng.random(n_envs) < p_done, info))
    return out


def _same_files(a: FakeH5, b: FakeH5):
    assert list(a.files) == list(b.files)
    for name in a.files:
        fa, fb = a.files[name], b.files[name]
        assert set(fa) == set(fb)
        assert fa["__attrs__"] == fb["__attrs__"], name
        for k in fa:
            if k != "__attrs__":
                assert fa[k].data.dtype == fb[k].data.dtype and fa[k].data.shape == fb[k].data.shape, (name, k)
                assert np.array_equal(fa[k].data, fb[k].data), (name, k)


@pytest.mark.parametrize("n_envs,n_steps,chunk,max_rows,p_done", [(5, 37, 8, 1000, 0.15), (3, 50, 64, 40, 0.2), (16, 25, 1, 90, 0.05),
                                                                 (4, 30, 7, 1000, 0.0), (2, 12, 5, 1000, 1.0)])
def test_recorder_files_equal_the_reference_loop(n_envs, n_steps, chunk, max_rows, p_done):
    extras = {"depth": {"shape": (3,), "dtype": np.float32}, "flag": {"shape": (), "dtype": np.float32}}
    steps = _steps(n_envs, n_steps, 6, 2, p_done, 11 * n_envs + n_steps, extras)
    ref, got = FakeH5(), FakeH5()
    reference_recorder_run(ref, "ds", n_envs, 6, 2, np.float32, np.float32, extras, max_rows, steps)
    with HDF5DataRecorder("ds", n_envs, _env(6, 2), extras, max_rows=max_rows, chunk_steps=chunk, backend=got) as rec:
        for obs, action, reward, done, info in steps:
            rec.append_to_buffer(torch.from_numpy(obs), torch.from_numpy(action), torch.from_numpy(reward), torch.from_numpy(done),
                                 {k: torch.from_numpy(v) for k, v in info.items()})
    _same_files(ref, got)
    assert sum(f["__attrs__"]["number_of_steps"] for f in got.files.values()) == n_envs * n_steps


def test_recorder_needs_h5py_like_the_reference_and_checks_its_arguments():
    with pytest.raises(AssertionError):
        HDF5DataRecorder("x.h5", 2, _env(3, 1), backend=FakeH5())
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            HDF5DataRecorder("x", 2, _env(3, 1))
    rec = HDF5DataRecorder("x", 2, _env(3, 1), backend=FakeH5())
    with pytest.raises(ValueError):
        rec.append_to_buffer(torch.zeros(3, 3), torch.zeros(3, 1), torch.zeros(3), torch.zeros(3, dtype=torch.bool), {})
    with pytest.raises(NotImplementedError):
        base = DataRecorderBase(1, chunk_steps=1)
        base.append_to_buffer(torch.zeros(1, 2), torch.zeros(1, 1), torch.zeros(1), torch.ones(1, dtype=torch.bool), {})


def test_sequential_collector_records_the_observation_the_action_came_from():
    """recorder/orbit.py:24-36 on an env that, like RoverEnv, rewrites its observation buffer in place."""
    class Env:
        def __init__(self):
            self.obs = torch.zeros(3, 4)
            self.t = 0

        def reset(self):
            return self.obs, {}

        def step(self, action):
            self.t += 1
            self.obs += 1.0  # in place
            done = torch.tensor([self.t % 2 == 0, False, self.t % 3 == 0])
            return self.obs, torch.full((3,), float(self.t)), done, torch.zeros(3, dtype=torch.bool), {}

    h5 = FakeH5()
    rec = HDF5DataRecorder("c", 3, _env(4, 1), max_rows=100, chunk_steps=4, backend=h5)
    SequentialCollector(Env(), None, rec, predict_fn=lambda m, o: o[:, :1] * 2.0, num_episodes=6).collect()
    rec.close()
    f = h5.files["c_0.h5"]
    assert f["__attrs__"]["number_of_steps"] == 18
    # every recorded action is twice the first column of the observation recorded beside it
    assert np.array_equal(f["actions"].data[:, 0], 2.0 * f["observations"].data[:, 0])
    assert sorted(f["rewards"].data[:, 0].tolist()) == sorted([float(t) for t in range(1, 7)] * 3)
