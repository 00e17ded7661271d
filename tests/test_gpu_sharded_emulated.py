"""Sharded update on ONE GPU: G contexts stand for G ranks and the three exchanges of
amclj_b200/sharded.py (max, totals, row rebalance) are done by hand, so the kernels and the
planner of the N > 1 path are checked by the single-GPU gpu tier.  (NCCL itself is exercised by
tools/sharded_check.py under torchrun on a multi-GPU box, and the choreography by the gloo
test.)  The global particle list must equal the single-context update bit for bit."""
import numpy as np
import pytest

from amclj_b200 import _lib, maps, workloads
from amclj_b200.sharded import slots_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n", [(2, 20_000), (3, 10_001), (8, 64_000)])
def test_sharded_equals_single(world, n):
    import torch

    g = maps.lgfloor()
    w = workloads.c3(n)
    s = w.scan
    ref = _lib.Context(0, "NS")
    ref.set_map(g)
    ref.set_particles(w.x, w.y, w.theta)
    shards = []
    for r in range(world):
        lo, cnt = slots_of(r, world, n)
        c = _lib.Context(0, "NS")
        c.set_map(g)
        c.set_shard(r, world, lo, n)
        c.set_particles(w.x[lo:lo + cnt], w.y[lo:lo + cnt], w.theta[lo:lo + cnt])
        c.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        shards.append((c, lo, cnt))
    for it in range(2):
        ref.filter_update(w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.u0, seed=9 + it)
        # 1. motion + sensor, all-reduce(max)
        for c, _, _ in shards:
            c.shard_motion_sensor_async(w.curr, w.prev, 9 + it)
            c.synchronize()
        mx = max(float(c.device_tensor(_lib.BUF_RAW_MAX, (1,), "<f4").cpu()[0]) for c, _, _ in shards)
        totals = []
        for c, _, _ in shards:
            c.device_tensor(_lib.BUF_RAW_MAX, (1,), "<f4").fill_(mx)
            torch.cuda.synchronize()
            c.shard_cdf_async()
            c.synchronize()
            totals.append(int(c.device_tensor(_lib.BUF_TOTAL, (1,), "<i8").cpu()[0]))
        # 2. plan from the all-gathered totals
        tot = np.array(totals, np.int64).view(np.uint64)
        r, off, m_lo, m_cnt, send = _lib.plan_systematic(tot, n, w.u0)
        T = int(tot.sum(dtype=np.uint64))
        # 3. generate + rebalance (all-to-all by hand)
        outs = []
        for (c, _, _), rk in zip(shards, range(world)):
            cnt = int(m_cnt[rk])
            c.reserve_stage(max(cnt, slots_of(rk, world, n)[1]))
            c.shard_generate_async(T, int(off[rk]), r, int(m_lo[rk]), cnt)
            c.synchronize()
            outs.append(c.device_tensor(_lib.BUF_STAGE_OUT, (cnt, 4), "<f4").clone() if cnt else torch.empty(0, 4, device="cuda"))
        for d, (c, lo, cnt) in enumerate(shards):
            parts = []
            for src in range(world):
                a = int(send[src, :d].sum())
                parts.append(outs[src][a:a + int(send[src, d])])
            inbox = torch.cat(parts, 0)
            assert inbox.shape[0] == cnt
            c.device_tensor(_lib.BUF_STAGE_IN, (cnt, 4), "<f4").copy_(inbox)
            torch.cuda.synchronize()
            c.shard_adopt_async(cnt)
            c.synchronize()
    rx, ry, rt, rw = ref.get_particles()
    got = [c.get_particles() for c, _, _ in shards]
    for k, refv in enumerate((rx, ry, rt)):
        assert np.array_equal(np.concatenate([g_[k] for g_ in got]), refv)
    np.testing.assert_allclose(np.concatenate([g_[3] for g_ in got]), rw, rtol=1e-7)
    for c, _, _ in shards:
        c.close()
    ref.close()


def test_weights_do_not_depend_on_task_split(monkeypatch):
    """The same particles give bit-identical raw weights whether a particle's beams run as one
    warp task or are split into chunks (fixed-point segment sums)."""
    g = maps.lgfloor()
    w = workloads.c3(5000)
    s = w.scan
    out = []
    for chunks in ("1", "3", "12"):
        monkeypatch.setenv("AMCLJ_CHUNKS", chunks)
        c = _lib.Context(0, "NS")
        c.set_map(g)
        c.set_particles(w.x, w.y, w.theta)
        c.sensor_update(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        out.append(c.raw_weights())
        c.close()
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])


@pytest.mark.parametrize("world,n", [(2, 20_000), (3, 10_001), (8, 64_000)])
def test_sharded_pull_equals_single(world, n):
    """Peer-memory variant (amclj_shard_pull_async): every rank fills its own slots by reading
    the owning rank's CDF and particles; here the 'peers' are contexts of one process."""
    import torch

    g = maps.lgfloor()
    w = workloads.c3(n)
    s = w.scan
    ref = _lib.Context(0, "NS")
    ref.set_map(g)
    ref.set_particles(w.x, w.y, w.theta)
    shards = []
    for r in range(world):
        lo, cnt = slots_of(r, world, n)
        c = _lib.Context(0, "NS")
        c.set_map(g)
        c.set_shard(r, world, lo, n)
        c.set_particles(w.x[lo:lo + cnt], w.y[lo:lo + cnt], w.theta[lo:lo + cnt])
        c.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        shards.append(c)
    for c in shards:
        for r, peer in enumerate(shards):
            c.shard_attach_local(r, peer)
    for it in range(3):
        ref.filter_update(w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.u0, seed=9 + it)
        for c in shards:
            c.shard_motion_sensor_async(w.curr, w.prev, 9 + it)
            c.synchronize()
        mx = max(float(c.device_tensor(_lib.BUF_RAW_MAX, (1,), "<f4").cpu()[0]) for c in shards)
        totals = []
        for c in shards:
            c.device_tensor(_lib.BUF_RAW_MAX, (1,), "<f4").fill_(mx)
            torch.cuda.synchronize()
            c.shard_cdf_async()
            c.synchronize()
            totals.append(int(c.device_tensor(_lib.BUF_TOTAL, (1,), "<i8").cpu()[0]))
        tot = torch.tensor(totals, dtype=torch.int64, device="cuda")
        for c in shards:
            c.device_tensor(_lib.BUF_TOTALS_ALL, (16,), "<i8")[:world].copy_(tot)
        torch.cuda.synchronize()
        for c in shards:
            c.shard_pull_async(w.u0)
            c.synchronize()
    rx, ry, rt, rw = ref.get_particles()
    got = [c.get_particles() for c in shards]
    for k, refv in enumerate((rx, ry, rt)):
        assert np.array_equal(np.concatenate([g_[k] for g_ in got]), refv)
    np.testing.assert_allclose(np.concatenate([g_[3] for g_ in got]), rw, rtol=1e-7)
    for c in shards:
        c.close()
    ref.close()


@pytest.mark.parametrize("world,n", [(1, 30_000), (2, 200_000), (3, 100_003), (8, 400_000)])
def test_multi_context_equals_single(world, n):
    """amclj_create_multi (one process, one caller thread; the shards emulated on GPU 0): the kernels
    exchange {max, total} through the peer-memory mailboxes and pull their slots; three updates and the
    host-buffer call give the particle list of ONE ctx holding everything, bit for bit."""
    g = maps.lgfloor()
    w = workloads.c3(n)
    s = w.scan
    with _lib.Context(0, "NS") as ref, _lib.MultiContext([0] * world, "NS") as m:
        ref.set_map(g)
        m.set_map(g)
        ref.set_particles(w.x, w.y, w.theta)
        m.set_particles(w.x, w.y, w.theta)
        m.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        for it in range(3):
            curr, prev = (0.2 * (it + 1), 0.0, 0.1 * (it + 1)), (0.2 * it, 0.0, 0.1 * it)
            ref.filter_update(curr, prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, 0.1 + 0.3 * it, seed=40 + it)
            m.filter_update_async(curr, prev, 40 + it, 0.1 + 0.3 * it)
            m.synchronize()
            for a, b in zip(m.get_particles()[:3], ref.get_particles()[:3]):
                assert np.array_equal(a, b)
        np.testing.assert_allclose(m.pose_estimate(), ref.pose_estimate(), rtol=1e-9)
        # host-buffer entry point
        bufs = [a.copy() for a in ref.get_particles()]
        mine = [a.copy() for a in bufs]
        ref.filter_update_host(*bufs, (0.9, 0.1, 0.4), (0.6, 0.0, 0.3), s.ranges, s.angle_min, s.angle_inc, s.range_max, 0.77, seed=5)
        m.filter_update_host(*mine, (0.9, 0.1, 0.4), (0.6, 0.0, 0.3), s.ranges, s.angle_min, s.angle_inc, s.range_max, 0.77, seed=5)
        for a, b in zip(mine[:3], bufs[:3]):
            assert np.array_equal(a, b)
        assert m.rank(0).stats().sensor_path == 3


def test_multi_context_skewed_weights():
    """Nearly all weight on the last shard: most slots are pulled from a peer."""
    g = maps.lgfloor()
    n, world = 120_000, 4
    x, y, th = workloads.uniform_particles(g, n, seed=61)
    gx, gy, gt = workloads.gaussian_particles(n // 4)
    x[-len(gx):], y[-len(gx):], th[-len(gx):] = gx, gy, gt  # the last shard sits on the true pose
    s = workloads.c2(10).scan
    with _lib.Context(0, "NS") as ref, _lib.MultiContext([0] * world, "NS") as m:
        for q in (ref, m):
            q.set_map(g)
            q.set_particles(x, y, th)
        m.stage_scan(s.ranges, s.angle_min, s.angle_inc, s.range_max)
        ref.filter_update((0, 0, 0), (0, 0, 0), s.ranges, s.angle_min, s.angle_inc, s.range_max, 0.5, seed=1)
        m.filter_update_async((0, 0, 0), (0, 0, 0), 1, 0.5)
        m.synchronize()
        for a, b in zip(m.get_particles()[:3], ref.get_particles()[:3]):
            assert np.array_equal(a, b)
        idx = ref.resample_indices()
        assert (idx >= n - len(gx)).mean() > 0.9  # the test does what it says


def test_rendezvous_alone_and_refused_on_a_shared_device():
    """amclj_shard_rendezvous_async: a shard that is its own world passes straight through (twice: the count
    advances); shards that share one GPU are refused (kernels that wait on one another must not share a device)."""
    g = maps.lgfloor()
    w = workloads.c3(4000)
    c = _lib.Context(0, "NS")
    c.set_map(g)
    c.set_shard(0, 1, 0, w.n)
    c.set_particles(w.x, w.y, w.theta)
    c.shard_attach_local(0, c)
    c.shard_rendezvous_async()
    c.shard_rendezvous_async()
    c.synchronize()
    c.close()
    shards = []
    for r in range(2):
        lo, cnt = slots_of(r, 2, w.n)
        d = _lib.Context(0, "NS")
        d.set_map(g)
        d.set_shard(r, 2, lo, w.n)
        d.set_particles(w.x[lo:lo + cnt], w.y[lo:lo + cnt], w.theta[lo:lo + cnt])
        shards.append(d)
    for d in shards:
        for r, peer in enumerate(shards):
            d.shard_attach_local(r, peer)
    with pytest.raises(_lib.AmcljError) as e:
        shards[0].shard_rendezvous_async()
    assert e.value.code == _lib.ERR_STATE
    for d in shards:
        d.close()
