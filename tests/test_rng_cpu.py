"""CPU: the counter-based variates of the reset path.  (i) the numpy restatement (oracle/philox.py) reproduces the
known-answer vectors Random123 publishes for Philox4x32-10 (kat_vectors); (ii) the library's host export of the very
functions the kernels run (csrc/rng.cuh, compiled for the host) agrees with the restatement bit for bit; (iii) the
spawn draw is a permutation (sampling without replacement, randomizations.py:22) and the variates are uniform."""
import ctypes as C

import numpy as np
import pytest

from isaac_rover_orbit_b200 import _lib
from oracle import philox as PX

# Random123 kat_vectors, "philox4x32 10": counter, key -> expected
KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(ctr, key, want):
    got = PX.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
    assert tuple(int(v) for v in got) == want
    lib = _lib.load()
    out = (C.c_uint32 * 4)()
    assert lib.rover_philox4x32_10(C.byref((C.c_uint32 * 4)(*ctr)), C.byref((C.c_uint32 * 2)(*key)), C.byref(out)) == 0
    assert tuple(out) == want


@pytest.mark.parametrize("seed,step,n,rounds,spawns", [(0, 0, 1, 1, 1), (7, 3, 257, 16, 514), (2 ** 40 + 5, 2 ** 33, 1000, 6, 1000),
                                                        (123, 9, 64, 21, 4096)])
def test_host_export_equals_restatement(seed, step, n, rounds, spawns):
    sp, yaw, head, theta = _lib.rng_variates(seed, step, n, rounds, spawns)
    sp2, yaw2, head2, theta2 = PX.variates(seed, step, n, rounds, spawns)
    assert np.array_equal(sp, sp2) and np.array_equal(yaw, yaw2) and np.array_equal(head, head2)
    assert np.array_equal(theta, theta2)
    assert yaw.dtype == np.float32 and (yaw >= 0).all() and (yaw < 1).all() and (theta < 1).all()


@pytest.mark.parametrize("spawns", [1, 2, 3, 31, 32, 33, 1000, 16384, 32768 + 1])
def test_spawn_draw_is_a_permutation(spawns):
    for step in (0, 1, 12345):
        sp = _lib.rng_variates(11, step, spawns, 1, spawns)[0]
        assert np.array_equal(np.sort(sp), np.arange(spawns)), "every row exactly once: distinct envs draw distinct rows"
    a = _lib.rng_variates(11, 0, spawns, 1, spawns)[0]
    b = _lib.rng_variates(11, 1, spawns, 1, spawns)[0]
    if spawns >= 31:
        assert not np.array_equal(a, b), "the permutation changes from step to step"


def test_variates_are_uniform_and_decorrelated():
    _, yaw, head, theta = _lib.rng_variates(5, 17, 65536, 8, 131072)
    for u in (yaw, head, theta.ravel()):
        assert abs(float(u.mean()) - 0.5) < 5e-3 and abs(float(u.var()) - 1 / 12) < 2e-3
    assert abs(float(np.corrcoef(yaw, head)[0, 1])) < 0.02
    assert abs(float(np.corrcoef(theta[:, 0], theta[:, 1])[0, 1])) < 0.02
    _, yaw2, _, _ = _lib.rng_variates(5, 18, 65536, 8, 131072)
    assert abs(float(np.corrcoef(yaw, yaw2)[0, 1])) < 0.02, "consecutive steps"
    # the rows of the first envs are spread over the table, not clustered
    sp = _lib.rng_variates(5, 17, 4096, 1, 131072)[0]
    assert abs(float(sp.mean()) / 131072 - 0.5) < 0.03
