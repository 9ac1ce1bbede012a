"""Golden call trace of the UNMODIFIED reference trainer (rover_envs/utils/skrl_utils.py) driven by the fake env /
agent of tests/trainer_fakes.py.  skrl itself is absent: its Trainer base class (1.1.0) is restated in
oracle/ref_loader.py.  Run in the build container (needs /root/reference):  python tests/golden/make_trainer_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from trainer_fakes import FakeAgent, FakeEnv  # noqa: E402


def trace(mode: str, timesteps: int, headless: bool):
    ref = ref_loader.load("skrl_utils")
    log = []
    trainer = ref.SkrlSequentialLogTrainer(env=FakeEnv(log), agents=FakeAgent(log),
                                           cfg={"timesteps": timesteps, "disable_progressbar": True, "headless": headless})
    getattr(trainer, mode)()
    return log


if __name__ == "__main__":
    out = {"train_5": trace("train", 5, True), "eval_4_headless": trace("eval", 4, True),
           "eval_3_render": trace("eval", 3, False)}
    with open(os.path.join(HERE, "trainer_calls.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print({k: len(v) for k, v in out.items()})
