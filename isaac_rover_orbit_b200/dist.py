"""Env sharding and the one collective on the path: the episode-statistics reduction.

The reference is single-process / single-GPU; every MDP term is per-env, so environments
shard across ranks with no data-path exchange.  The only cross-env quantities are the logging
statistics the ORBIT managers' ``reset()`` produce (``extras["log"]``, consumed at rover_envs/utils/skrl_utils.py:
139-142): per-term episodic reward means, termination counts and command metrics over the envs that reset.
Each rank's fused post-step kernel accumulates the SUMS and the COUNT into one 16-float vector; ranks combine
them with a single ``all_reduce(SUM)`` (NCCL over NVLink on GPUs, gloo in the CPU tests) and divide afterwards
-- sum / count, not a mean of means.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from .config import REWARD_TERMS, TERMINATION_TERMS

STATS_LEN = 16
IDX_ERR_POS, IDX_ERR_HEADING, IDX_NUM_RESETS, IDX_EXHAUSTED, IDX_TIME_RESAMPLES = 11, 12, 13, 14, 15


def shard_range(num_envs: int, rank: int, world: int) -> tuple:
    """Block partition of the env axis: rank r owns [r*N/G, (r+1)*N/G) (remainder spread over the first ranks)."""
    base, rem = divmod(num_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def spawn_rows(num_envs: int, rank: int, world: int) -> tuple:
    """Rows of the global spawn table (``2 * num_envs`` rows) a rank draws from: twice its env range, so the slices are
    disjoint, cover the table and match ``shard_range`` also when ``num_envs`` does not divide evenly -- the
    without-replacement draw of mdp/randomizations.py:22 stays shard-local."""
    lo, hi = shard_range(num_envs, rank, world)
    return 2 * lo, 2 * hi


def episode_log(stats: torch.Tensor, episode_length_s: float = 150.0) -> dict:
    """``extras["log"]`` from a (globally reduced) statistics vector, as the ORBIT managers build it."""
    s = stats.detach().double().cpu()
    k = max(float(s[IDX_NUM_RESETS]), 1.0)
    log = {f"Episode Reward/{name}": float(s[i]) / k / episode_length_s for i, name in enumerate(REWARD_TERMS)}
    log.update({f"Episode Termination/{name}": int(round(float(s[7 + i]))) for i, name in enumerate(TERMINATION_TERMS)})
    log["Metrics/target_pose/error_pos"] = float(s[IDX_ERR_POS]) / k
    log["Metrics/target_pose/error_heading"] = float(s[IDX_ERR_HEADING]) / k
    log["num_resets"] = int(round(float(s[IDX_NUM_RESETS])))
    return log


class EpisodeStats:
    """Owns the hand-off of the per-rank statistics vector to the collective.

    ``all_reduce_async`` snapshots ``buf.stats`` into a staging vector, clears the accumulator for the next
    interval and starts the all-reduce without blocking the compute stream; ``result`` waits for it.
    """

    def __init__(self, buf, world_size: int | None = None, group=None):
        self.stats = buf.stats if hasattr(buf, "stats") else buf
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.staging = torch.zeros_like(self.stats)
        self._work = None

    def all_reduce_async(self):
        if self._work is not None:
            self._work.wait()
        self.staging.copy_(self.stats)
        self.stats.zero_()
        if self.world > 1:
            self._work = dist.all_reduce(self.staging, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return self

    def result(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self.staging

    def log(self, episode_length_s: float = 150.0) -> dict:
        return episode_log(self.result(), episode_length_s)


class P2PStats:
    """Episode statistics across the GPUs of one node without a collective on the step path.

    Every rank owns a mailbox (one 256-byte slot per rank, ``cudaMalloc``'ed and exported through CUDA IPC); the handles
    are exchanged once with ``all_gather`` and every rank maps every mailbox.  Passing the object to
    ``ops.mdp_post_step(..., xchg=...)`` makes the last block of that launch add the step's statistics to the rank's
    running totals (fp64) and store them into its slot of every mailbox (peer stores over NVLink, sequence-locked).
    ``read()`` sums the slots of the local mailbox in rank order -- the global running totals as of each rank's latest
    published step -- and ``interval()`` returns the difference to the previous call (the numbers ``episode_log``
    divides: sum / count over all ranks, not a mean of means).  ``EpisodeStats`` (one all-reduce) remains the portable
    path (several nodes, CPU tests)."""

    def __init__(self, device, rank: int | None = None, world: int | None = None, group=None):
        from . import _lib

        self._lib = _lib
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("P2PStats needs a CUDA device; there is no CPU fallback (use EpisodeStats)")
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        lib = _lib.load()
        with torch.cuda.device(self.device):
            mb = C.c_void_p()
            _lib.check(lib.rover_p2p_alloc(C.byref(mb), _lib.MAILBOX_SLOT_BYTES * self.world))
            self._mailbox = mb.value
            handle = (C.c_uint8 * 64)()
            _lib.check(lib.rover_p2p_export(C.c_void_p(self._mailbox), C.byref(handle)))
            # the 64-byte IPC handles are the only thing the process group ever carries (once, here); under gloo -- the
            # one-GPU test of the protocol: two processes sharing a device -- they travel as host tensors
            on_host = self.world > 1 and dist.get_backend(group) == "gloo"
            mine = torch.tensor(list(handle), dtype=torch.uint8, device="cpu" if on_host else self.device)
            if self.world > 1:
                gathered = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(gathered, mine, group=group)
            else:
                gathered = [mine]
            ptrs = []
            self._opened = []
            for r, h in enumerate(gathered):
                if r == self.rank:
                    ptrs.append(self._mailbox)
                    continue
                raw = (C.c_uint8 * 64)(*h.cpu().tolist())
                p = C.c_void_p()
                _lib.check(lib.rover_p2p_open(C.byref(raw), C.byref(p)))
                ptrs.append(p.value)
                self._opened.append(p.value)
            self._peer_table = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
            self._cumulative = torch.zeros(STATS_LEN, dtype=torch.float64, device=self.device)
            self._sequence = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._out = torch.zeros(STATS_LEN, dtype=torch.float64, device=self.device)
            self._last = torch.zeros(STATS_LEN, dtype=torch.float64, device=self.device)
        self.struct = _lib.StatsExchange(self._peer_table.data_ptr(), self._cumulative.data_ptr(),
                                         self._sequence.data_ptr(), self.rank, self.world)
        from . import torch_ops

        self.desc = torch_ops.descriptor(self.struct)  # argument of torch.ops.rover_b200.mdp_post_step(..., xchg=)
        # the local mailbox as a tensor (not torch-allocated memory: cudaMalloc'ed so that it can be exported over IPC)
        self._mailbox_words = _lib.MAILBOX_SLOT_BYTES * self.world // 8
        if self.world > 1:
            dist.barrier(group=group)  # every mailbox is mapped everywhere before anyone publishes

    def flush(self) -> None:
        """Publish this rank's running totals now (``rover_stats_publish``).  The single-launch step with in-kernel
        variates publishes at the START of a launch the totals of the launches before it (off the step's critical path),
        so the mailboxes run one launch behind; for exact totals every rank calls ``flush()``, synchronises, and the
        ranks meet at a barrier before ``read()`` (``totals()`` does all of that)."""
        self._lib.check(self._lib.load().rover_stats_publish(C.byref(self.struct), self._lib.current_stream(self.device)))

    def totals(self, group=None) -> torch.Tensor:
        """Exact global running totals ``[16]`` f64 as of the last launch of EVERY rank (collective: flush, barrier, read)."""
        self.flush()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=group)
        return self.read()

    def local_totals(self) -> torch.Tensor:
        """This rank's own running totals ``[16]`` f64."""
        return self._cumulative

    def read(self) -> torch.Tensor:
        """Global running totals ``[16]`` f64 (device tensor, enqueued on the current stream) as of each rank's latest
        PUBLISHED launch (see ``flush``)."""
        self._lib.check(self._lib.load().rover_stats_read(C.c_void_p(self._mailbox), self.world,
                                                          C.c_void_p(self._out.data_ptr()),
                                                          self._lib.current_stream(self.device)))
        return self._out  # (the mailbox is raw cudaMalloc memory, so this one call stays on the plain C ABI)

    def interval(self) -> torch.Tensor:
        """Totals accumulated since the previous ``interval()`` call (what ``episode_log`` expects)."""
        now = self.read().clone()
        delta = now - self._last
        self._last = now
        return delta

    def log(self, episode_length_s: float = 150.0) -> dict:
        return episode_log(self.interval(), episode_length_s)

    def close(self) -> None:
        lib = self._lib.load()
        for p in self._opened:
            lib.rover_p2p_close(C.c_void_p(p))
        self._opened = []
        if self._mailbox:
            lib.rover_p2p_free(C.c_void_p(self._mailbox))
            self._mailbox = None
