"""Multi-GPU plumbing: environments shard by batch, one process per GPU, no data-path collective.

The only exchange of the path is the end-of-episode gather of the per-environment metric matrix
(the reference itself is single-process).  Works with NCCL (CUDA tensors) and gloo (CPU tensors,
used by the world-size-2 CPU tests).
"""
import torch
import torch.distributed as dist

METRIC_COLUMNS = ("success", "timeout", "n_steps", "nav_time", "n_collisions", "n_wall_collisions", "n_frozen",
                  "n_too_close", "min_dist")


def shard_range(total_envs, rank, world):
    """Contiguous shard [lo, hi) of the global environment ids owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_case_ids(total_envs, rank, world, test_size=500):
    """Reference test-case ids (seed = 1000 + case) of this rank's environments: global env id modulo test_size, so
    results do not depend on the number of ranks."""
    lo, hi = shard_range(total_envs, rank, world)
    return [g % test_size for g in range(lo, hi)]


class EpisodeMetrics:
    """Per-environment episode counters (METRIC_COLUMNS; what simple_test.py accumulates from `info` and pickles,
    simple_test.py:232-258, 306-319), updated from the flag word of every step.  On a CUDA device the update is one libsnb launch
    (snb_episode_metrics_update); CPU tensors (the gloo tests of the sharding / gather logic) take the same arithmetic in torch."""

    def __init__(self, n_envs, device, time_step):
        self.m = torch.zeros(n_envs, len(METRIC_COLUMNS), dtype=torch.float64, device=device)
        self.m[:, 8] = float("inf")
        self.dt = float(time_step)
        self.live = torch.ones(n_envs, dtype=torch.uint8, device=device)

    def update(self, flags, dmin, stream=None):
        if self.m.is_cuda:
            from . import _capi
            assert flags.dtype == torch.int32 and dmin.dtype == torch.float64 and flags.is_contiguous() and dmin.is_contiguous()
            _capi.check(_capi.lib.snb_episode_metrics_update(_capi.ptr(self.m), _capi.ptr(self.live), _capi.ptr(flags), _capi.ptr(dmin),
                                                             self.dt, self.m.shape[0], _capi.stream_ptr(stream)), "snb_episode_metrics_update")
            return
        F = flags.to(torch.int64)
        live = self.live.bool()
        f = live.to(torch.float64)
        self.m[:, 2] += f
        self.m[:, 3] += f * self.dt
        self.m[:, 4] += f * ((F & 4) != 0)
        self.m[:, 5] += f * ((F & 8) != 0)
        self.m[:, 6] += f * ((F & 16) != 0)
        self.m[:, 7] += f * ((F & 32) != 0)
        self.m[:, 8] = torch.where(live, torch.minimum(self.m[:, 8], dmin.to(torch.float64)), self.m[:, 8])
        self.m[:, 0] = torch.where(live & ((F & 1) != 0), torch.ones_like(self.m[:, 0]), self.m[:, 0])
        self.m[:, 1] = torch.where(live & ((F & 2) != 0), torch.ones_like(self.m[:, 1]), self.m[:, 1])
        self.live = (live & ((F & 64) == 0)).to(torch.uint8)


def gather_metrics(local, total_envs=None, group=None):
    """all_gather of the [B_local, 9] metric matrix -> [B_total, 9] in global env order on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if total_envs is None:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous(), group=group)
        return torch.cat(parts, 0)
    sizes = [shard_range(total_envs, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, local.shape[1], dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)


def summarize(metrics):
    """Rates over all environments (what simple_test.py pickles per episode, aggregated)."""
    m = metrics.double()
    n = max(1, m.shape[0])
    return dict(episodes=int(m.shape[0]), success_rate=float(m[:, 0].sum() / n), timeout_rate=float(m[:, 1].sum() / n),
                collision_rate=float((m[:, 4] > 0).sum() / n), mean_steps=float(m[:, 2].mean()), mean_nav_time=float(m[:, 3].mean()),
                collisions_per_episode=float(m[:, 4].mean()), wall_collisions_per_episode=float(m[:, 5].mean()),
                frozen_steps_per_episode=float(m[:, 6].mean()), too_close_steps_per_episode=float(m[:, 7].mean()),
                min_dist=float(m[:, 8].min()) if m.shape[0] else float("inf"))
