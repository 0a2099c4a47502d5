"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding and the end-of-episode metric gather."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _flags_for(env_ids, step):
    """Deterministic synthetic flag words / dmin per global env id (stands in for snb_env_step outputs)."""
    g = torch.tensor(env_ids, dtype=torch.int64)
    f = torch.zeros_like(g)
    f |= ((g + step) % 7 == 0).long() * 4
    f |= ((g * 3 + step) % 11 == 0).long() * 8
    f |= ((g + 2 * step) % 5 == 0).long() * 16
    f |= ((g + step) % 3 == 0).long() * 32
    done = (step >= 3 + g % 6)
    f |= done.long() * 64 | (done & (g % 2 == 0)).long() * 1 | (done & (g % 2 == 1)).long() * 2
    dmin = 0.1 + ((g * 7 + step * 3) % 13).float() / 10
    return f.to(torch.int32), dmin.double()


def _episode(env_ids):
    sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
    from snb.dist import EpisodeMetrics
    em = EpisodeMetrics(len(env_ids), "cpu", 0.25)
    for step in range(12):
        f, d = _flags_for(env_ids, step)
        em.update(f, d)
    return em.m


def _worker(rank, world, port, total, q):
    sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
    from snb.dist import gather_metrics, shard_range, summarize
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    local = _episode(list(range(lo, hi)))
    allm = gather_metrics(local, total_envs=total)
    q.put((rank, allm.clone(), summarize(allm)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
    from snb.dist import global_case_ids, shard_range
    for total in (0, 1, 7, 8192, 1001):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    ids8 = sum((global_case_ids(1000, k, 8) for k in range(8)), [])
    assert ids8 == [g % 500 for g in range(1000)]        # independent of the number of ranks


@pytest.mark.timeout(120)
def test_metric_gather_world2_equals_single_process():
    total, world = 37, 2          # ragged shards: 19 + 18
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    single = _episode(list(range(total)))
    for rank, allm, summ in res:
        assert torch.equal(allm, single), rank
        assert 0.0 <= summ["success_rate"] <= 1.0 and summ["mean_steps"] > 0


def _oracle_episode_metrics(cases):
    """Real flag words: the CPU oracle (oracle/rollout_oracle.py) runs the ORCA episodes of `cases`; EpisodeMetrics consumes its per-step
    flags / dmin exactly as snb/rollout.py feeds it the kernel's."""
    for p in (os.path.join(ROOT, "safe-interactive-crowdnav_b200"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    import rollout_oracle as RO
    from snb.dist import EpisodeMetrics
    env, pcfg, dcfg, rcfg = RO.reset(cases, 5, time_limit=8.0, n_threads=1)
    m_ref, trace = RO.run_episodes(env, pcfg, dcfg, rcfg, n_threads=1)
    em = EpisodeMetrics(len(cases), "cpu", 0.25)
    for flags, dmin, live in trace:
        em.update(torch.from_numpy(np.where(live, flags, 0).astype(np.int32)), torch.from_numpy(dmin))
    assert np.array_equal(em.m.numpy(), m_ref)
    return em.m


def _worker_real(rank, world, port, total, q):
    sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
    from snb.dist import gather_metrics, global_case_ids, summarize
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    local = _oracle_episode_metrics(global_case_ids(total, rank, world, test_size=500))
    allm = gather_metrics(local, total_envs=total)
    q.put((rank, allm.clone(), summarize(allm)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_episode_metrics_world2_on_oracle_generated_flags():
    """configs[4] host logic on REAL episodes: 25 ORCA test cases sharded 13 + 12 over two gloo ranks; the gathered metric matrix and
    the summary equal a single-process run of all 25 cases (results do not depend on the number of ranks)."""
    total, world = 25, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_real, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    single = _oracle_episode_metrics(list(range(total)))
    for rank, allm, summ in res:
        assert torch.equal(allm, single), rank
        assert summ["episodes"] == total and summ["success_rate"] + summ["timeout_rate"] == 1.0
