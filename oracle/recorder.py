"""TEST INFRASTRUCTURE (never imported by the product package): the reference's recorder restated step by step --
rover_envs/utils/recorder/data_recorder/base.py:42-67 (per-rover Python lists, written when the rover reports done) and
hdf_recorder.py:34-88 (file layout, roll-over at max_rows, number_of_steps, flush order, truncation) -- against an
in-memory stand-in for the small part of h5py's API the reference uses.  tests/test_recorder_cpu.py checks
isaac_rover_orbit_b200/recorder.py (device staging ring, batched transfer, regrouping by done events) against it."""
import numpy as np


class FakeH5:
    """`h5py`-shaped module: File(name, mode) context manager, create_dataset, dataset slicing / resize, attrs."""

    def __init__(self):
        self.files = {}
        self.opens = []

    class _Dataset:
        def __init__(self, shape, dtype):
            self.data = np.zeros(shape, dtype=dtype)

        def __setitem__(self, idx, value):
            self.data[idx] = np.asarray(value).reshape(self.data[idx].shape)

        def __getitem__(self, idx):
            return self.data[idx]

        def resize(self, size, axis=0):
            assert axis == 0
            self.data = self.data[:size].copy()

        @property
        def shape(self):
            return self.data.shape

    class _File:
        def __init__(self, store):
            self._store = store
            self.attrs = store.setdefault("__attrs__", {})

        def create_dataset(self, name, shape, dtype=None):
            self._store[name] = FakeH5._Dataset(shape, dtype)
            return self._store[name]

        def __getitem__(self, name):
            return self._store[name]

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def File(self, name, mode):
        self.opens.append((name, mode))
        if mode == "w":
            self.files[name] = {}
        return FakeH5._File(self.files[name])


def reference_recorder_run(h5, base_filename, num_envs, num_obs, num_actions, obs_dtype, act_dtype, extras, max_rows, steps,
                           close=True):
    """`steps`: list of (obs [N,num_obs], action [N,num_actions], reward [N], done [N] bool, info dict of [N,...] arrays).
    Returns nothing; the files are in ``h5.files``."""
    state = {"row": 0, "index": 0, "name": None}
    keys = ["observations", "actions", "rewards", "terminated"] + list(extras.keys())

    def new_file():  # hdf_recorder.py:34-50
        state["name"] = f"{base_filename}_{state['index']}.h5"
        state["index"] += 1
        with h5.File(state["name"], "w") as f:
            f.create_dataset("observations", (max_rows, num_obs), dtype=obs_dtype)
            f.create_dataset("actions", (max_rows, num_actions), dtype=act_dtype)
            f.create_dataset("rewards", (max_rows, 1), dtype=np.float32)
            f.create_dataset("terminated", (max_rows, 1), dtype=bool)
            for key, param in extras.items():
                f.create_dataset(key, (max_rows, *param["shape"]), dtype=param["dtype"])
            f.attrs["number_of_steps"] = 0

    def init_buffer():  # base.py:26-35
        return {k: [] for k in keys}

    def write_to_disk(rover):  # hdf_recorder.py:52-67
        chunk = {k: np.array(v) for k, v in buffers[rover].items()}
        n = len(chunk["observations"])
        nxt = state["row"] + n
        if nxt > max_rows:
            new_file()
            state["row"] = 0
            nxt = n
        with h5.File(state["name"], "a") as f:
            for k, v in chunk.items():
                f[k][state["row"]:nxt] = v
            f.attrs["number_of_steps"] += n
        state["row"] = nxt
        buffers[rover] = init_buffer()

    new_file()
    buffers = {r: init_buffer() for r in range(num_envs)}
    for obs, action, reward, done, info in steps:  # base.py:42-67
        for r in range(num_envs):
            buffers[r]["observations"].append(obs[r])
            buffers[r]["actions"].append(action[r])
            buffers[r]["rewards"].append(reward[r])
            buffers[r]["terminated"].append(done[r])
            for k in extras:
                buffers[r][k].append(info[k][r])
            if done[r]:
                write_to_disk(r)
                buffers[r] = init_buffer()
    if close:
        for r in range(num_envs):  # hdf_recorder.py:69-73
            if len(buffers[r]["observations"]) > 0:
                write_to_disk(r)
        with h5.File(state["name"], "a") as f:  # hdf_recorder.py:75-80
            for k in keys:
                f[k].resize(f.attrs["number_of_steps"], axis=0)
