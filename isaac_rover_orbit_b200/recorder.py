"""Offline-dataset recorder: rover_envs/utils/recorder/data_recorder/base.py:7-84 (``DataRecorderBase``),
hdf_recorder.py:10-88 (``HDF5DataRecorder``) and the collection loop recorder/orbit.py:9-38 (``SequentialCollectorOrbit``).

The reference copies the whole batch to the host every step (``.cpu().numpy()`` on obs / action / reward / done, base.py:69-77)
and then loops over the environments in Python.  Here a step is a handful of device-to-device copies into a staging ring
``[chunk, N, D]`` -- no host synchronisation on the step path -- and every ``chunk`` steps (and at ``flush``) the ring goes to
page-locked host memory in one transfer per dataset, where the episodes are regrouped with one pass over the ``done`` events.
What lands in the file is identical to the reference's output, row for row: an episode is written when its env reports
``done``, episodes that end on the same step are written in env order (base.py:56-67), ``flush`` writes the unfinished
episodes in env order (hdf_recorder.py:69-73), files roll over at ``max_rows`` (hdf_recorder.py:56-59) and carry the
``number_of_steps`` attribute.

File format: the reference writes HDF5 through ``h5py``.  ``HDF5DataRecorder`` does the same when ``h5py`` can be imported and
raises ``ImportError`` when it cannot (this build image has neither h5py nor PyTables, so the HDF5 bytes themselves are
untested here -- the call sequence into the h5py API is, against the reference's own, with an in-memory stand-in:
tests/test_recorder_cpu.py).  ``backend=`` accepts any module with ``File(name, mode)`` objects offering
``create_dataset`` / ``__getitem__`` / ``attrs`` like h5py's."""
from __future__ import annotations

import importlib
from typing import Any, Dict, Optional

import numpy as np
import torch

_CORE = ("observations", "actions", "rewards", "terminated")


class DataRecorderBase:
    """base.py:7-84.  ``append_to_buffer(obs, action, reward, done, info)`` per step; derived classes implement
    ``write_rows(rows)`` (one finished episode: ``{dataset: ndarray [len, ...]}``)."""

    def __init__(self, num_envs: int, extras: Optional[Dict[str, Dict[str, Any]]] = None, chunk_steps: int = 64) -> None:
        self.num_envs = int(num_envs)
        self.extras = extras or {}
        self.chunk_steps = int(chunk_steps)
        self._keys = _CORE + tuple(self.extras.keys())
        self._ring: Dict[str, torch.Tensor] | None = None   # device staging [chunk, N, ...]
        self._host: Dict[str, torch.Tensor] | None = None   # page-locked mirror
        self._fill = 0
        # unfinished episodes carried across chunks: per env, a list of row blocks per dataset
        self._carry = [{k: [] for k in self._keys} for _ in range(self.num_envs)]

    # ---- step path: device-to-device copies only
    def append_to_buffer(self, obs, action, reward, done, info: Dict[str, Any] | None = None) -> None:
        step = {"observations": obs, "actions": action, "rewards": reward, "terminated": done}
        for key in self.extras:
            step[key] = info[key]
        step = {k: torch.as_tensor(v) for k, v in step.items()}
        if self._ring is None:
            self._allocate(step)
        for k, v in step.items():
            dst = self._ring[k][self._fill]
            dst.copy_(v.reshape(dst.shape), non_blocking=True)
        self._fill += 1
        if self._fill == self.chunk_steps:
            self._drain()

    def _allocate(self, step: Dict[str, torch.Tensor]) -> None:
        self._ring, self._host = {}, {}
        for k, v in step.items():
            if v.shape[0] != self.num_envs:
                raise ValueError(f"recorder: {k} has {v.shape[0]} rows, expected num_envs = {self.num_envs}")
            tail = tuple(v.shape[1:]) if v.dim() > 1 else (1,) if k in ("rewards", "terminated") else ()
            shape = (self.chunk_steps, self.num_envs) + tail
            self._ring[k] = torch.empty(shape, dtype=v.dtype, device=v.device)
            host = torch.empty(shape, dtype=v.dtype)
            self._host[k] = host.pin_memory() if v.is_cuda else host

    # ---- every chunk_steps steps: one transfer per dataset, then the episodes in the reference's order
    def _drain(self) -> None:
        k_steps = self._fill
        if k_steps == 0:
            return
        for k in self._keys:
            self._host[k][:k_steps].copy_(self._ring[k][:k_steps], non_blocking=True)
        dev = self._ring["observations"].device
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()
        chunk = {k: self._host[k][:k_steps].numpy() for k in self._keys}
        done = chunk["terminated"].reshape(k_steps, self.num_envs).astype(bool)
        start = np.zeros(self.num_envs, dtype=np.int64)
        for t, n in zip(*np.nonzero(done)):  # row-major: by step, then by env -- the order of base.py:56-67
            rows = {}
            for k in self._keys:
                blocks = self._carry[n][k] + [chunk[k][start[n]:t + 1, n]]
                rows[k] = np.concatenate(blocks, axis=0) if len(blocks) > 1 else blocks[0]
                self._carry[n][k] = []
            start[n] = t + 1
            self.write_rows(rows)
        for n in range(self.num_envs):  # what is left of the chunk belongs to episodes still running
            if start[n] < k_steps:
                for k in self._keys:
                    self._carry[n][k].append(chunk[k][start[n]:k_steps, n].copy())
        self._fill = 0

    def flush(self) -> None:
        """hdf_recorder.py:69-73: the unfinished episodes, in env order."""
        self._drain()
        for n in range(self.num_envs):
            if self._carry[n]["observations"]:
                rows = {k: np.concatenate(self._carry[n][k], axis=0) for k in self._keys}
                self._carry[n] = {k: [] for k in self._keys}
                self.write_rows(rows)

    def write_rows(self, rows: Dict[str, np.ndarray]) -> None:
        raise NotImplementedError("write_rows is implemented by a derived class")


class HDF5DataRecorder(DataRecorderBase):
    """hdf_recorder.py:10-88: datasets ``observations [max_rows, num_obs]``, ``actions [max_rows, num_actions]``,
    ``rewards [max_rows, 1]`` (float32), ``terminated [max_rows, 1]`` (bool) + one per ``extras`` entry, the file attribute
    ``number_of_steps``; files ``<base>_<k>.h5`` rolling over when an episode does not fit any more."""

    def __init__(self, base_filename: str, num_envs: int, env, extras: Optional[Dict[str, Dict[str, Any]]] = None,
                 max_rows: int = 500_000, chunk_steps: int = 64, backend=None) -> None:
        super().__init__(num_envs, extras, chunk_steps)
        if base_filename.endswith(".h5") or base_filename.endswith(".hdf5"):
            raise AssertionError("Base filename should not end with .h5 or .hdf5")  # hdf_recorder.py:21-22
        if backend is None:
            try:
                backend = importlib.import_module("h5py")
            except ImportError as e:
                raise ImportError("HDF5DataRecorder needs h5py (as the reference does); pass backend= to write through "
                                  "another module with h5py's File / create_dataset / attrs interface") from e
        self._h5 = backend
        self.base_filename = base_filename
        self.max_rows = int(max_rows)
        self.current_row = 0
        self.current_file_index = 0
        self.num_observations = env.observation_space.shape[0]
        space = env.action_space
        self.num_actions = space.shape[0] if getattr(space, "shape", None) else 1  # Discrete -> 1 (hdf_recorder.py:28)
        self.env = env
        self._create_new_file()

    def _create_new_file(self) -> None:
        self.file_name = f"{self.base_filename}_{self.current_file_index}.h5"
        self.current_file_index += 1
        with self._h5.File(self.file_name, "w") as file:
            file.create_dataset("observations", (self.max_rows, self.num_observations), dtype=self.env.observation_space.dtype)
            file.create_dataset("actions", (self.max_rows, self.num_actions), dtype=self.env.action_space.dtype)
            file.create_dataset("rewards", (self.max_rows, 1), dtype=np.float32)
            file.create_dataset("terminated", (self.max_rows, 1), dtype=bool)
            for key, param in self.extras.items():
                file.create_dataset(key, (self.max_rows, *param["shape"]), dtype=param["dtype"])
            file.attrs["number_of_steps"] = 0

    def write_rows(self, rows: Dict[str, np.ndarray]) -> None:
        """hdf_recorder.py:52-67 for one episode."""
        n = len(rows["observations"])
        next_index = self.current_row + n
        if next_index > self.max_rows:
            self._create_new_file()
            self.current_row = 0
            next_index = n
        with self._h5.File(self.file_name, "a") as file:
            for key, value in rows.items():
                file[key][self.current_row:next_index] = value
            file.attrs["number_of_steps"] += n
        self.current_row = next_index

    def _truncate_datasets(self) -> None:
        with self._h5.File(self.file_name, "a") as file:
            for key in self._keys:
                file[key].resize(file.attrs["number_of_steps"], axis=0)

    def close(self) -> None:
        self.flush()
        self._truncate_datasets()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()


class SequentialCollector:
    """recorder/orbit.py:9-38 (``SequentialCollectorOrbit``): ``num_episodes`` steps of ``predict_fn(model, obs)`` ->
    ``env.step`` -> ``recorder.append_to_buffer``.  The reference resets the whole env when ANY env is done (a host sync per
    step); ``RoverEnv`` resets finished envs inside its step, so the loop has nothing to wait for."""

    def __init__(self, env, model, recorder: DataRecorderBase, predict_fn=None, num_episodes: int = 1000):
        self.env, self.model, self.recorder = env, model, recorder
        self.predict_fn = predict_fn
        self.num_episodes = int(num_episodes)

    def collect(self) -> None:
        with torch.no_grad():
            obs, info = self.env.reset()
            for _ in range(self.num_episodes):
                action = self.predict_fn(self.model, obs)
                # the recorded observation is the one the action was computed from; RoverEnv.step rewrites its observation
                # buffer in place, so that row block is staged before the step (the reference's env returns a new tensor)
                before = obs.clone() if isinstance(obs, torch.Tensor) else obs
                next_obs, reward, done, truncated, info = self.env.step(action)
                self.recorder.append_to_buffer(before, action, reward, done, info)  # orbit.py:31: `done` = terminated
                obs = next_obs
