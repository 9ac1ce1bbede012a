"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of ``oracle/raycast.c`` (see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_raycast.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile ``raycast.c`` with the committed Makefile (gcc + OpenMP)."""
    src = os.path.join(_HERE, "raycast.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B" if force else "-s"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.rc_mesh_create.restype = ctypes.c_void_p
        lib.rc_mesh_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        lib.rc_mesh_destroy.argtypes = [ctypes.c_void_p]
        lib.rc_raycast.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.rc_num_nodes.restype = ctypes.c_int
        lib.rc_num_nodes.argtypes = [ctypes.c_void_p]
        _lib = lib
    return _lib


class Mesh:
    """Triangle mesh + BVH; ``raycast`` restates ORBIT ``raycast_mesh`` over warp ``mesh_query_ray``."""

    def __init__(self, vertices, faces):
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))
        self.faces = np.ascontiguousarray(np.asarray(faces, dtype=np.int32).reshape(-1, 3))
        lib = _load()
        self._h = lib.rc_mesh_create(self.vertices.ctypes.data, len(self.vertices), self.faces.ctypes.data,
                                     len(self.faces))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rc_mesh_destroy(self._h)
            self._h = None

    def raycast(self, starts: torch.Tensor, dirs: torch.Tensor, max_dist: float, brute: bool = False,
                return_t: bool = False):
        s = np.ascontiguousarray(starts.detach().cpu().numpy().astype(np.float32).reshape(-1, 3))
        d = np.ascontiguousarray(dirs.detach().cpu().numpy().astype(np.float32).reshape(-1, 3))
        n = s.shape[0]
        hits = np.empty((n, 3), dtype=np.float32)
        t = np.empty(n, dtype=np.float32)
        face = np.empty(n, dtype=np.int32)
        _load().rc_raycast(self._h, s.ctypes.data, d.ctypes.data, n, ctypes.c_float(max_dist), int(brute),
                           hits.ctypes.data, t.ctypes.data, face.ctypes.data)
        if return_t:
            return torch.from_numpy(hits), torch.from_numpy(t), torch.from_numpy(face)
        return torch.from_numpy(hits)
