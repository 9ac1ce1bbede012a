"""GPU: programmatic dependent launch (csrc/common.cuh) must not change a bit.  The step kernels may start while the kernel
in front of them retires; a closed loop (MDP step -> height scan -> policy + value -> Gaussian act, no synchronisation
inside) is run in two processes, ROVER_PDL=1 and ROVER_PDL=0 (the switch is read once per process), and the trajectory
checksums are compared."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(pdl: str) -> str:
    env = dict(os.environ, ROVER_PDL=pdl)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "soak_pdl.py"), "60", "3000"], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("TRAJECTORY")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    return lines[-1].split()[1]


def test_closed_loop_trajectory_is_the_same_with_and_without_pdl(cuda_device):
    assert _run("1") == _run("0")
