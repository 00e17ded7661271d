"""The N > 1 choreography (amclj_b200/sharded.py) on CPU: two gloo ranks, a stand-in backend
whose device work is done by the oracle.  The rebalanced global particle list must equal the
single-list systematic resampler's."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["AMCLJ_ROOT"])
import oracle
from amclj_b200 import _lib
from amclj_b200.sharded import ShardedFilter, slots_of

class OracleShard:
    """Backend protocol of sharded.CudaShard with the oracle standing in for the kernels."""
    def __init__(self, rows, raw):
        self.rows, self.raw = rows.copy(), raw.copy()
    def set_shard(self, rank, world, offset, n_global): self.n_global = n_global
    def reserve(self, rows): pass
    def empty(self, *shape, dtype): return torch.empty(*shape, dtype=dtype)
    def motion_sensor(self, curr, prev, seed): pass
    def raw_max(self):
        self.mx = torch.tensor([self.raw.max()], dtype=torch.float32); return self.mx
    def cdf(self):
        w = np.zeros(len(self.raw), np.float32)
        oracle.lib().amclj_oracle_linear_weights(self.raw, len(self.raw), float(self.mx[0]), w)
        self.q, tot = oracle.fixed_weights(w, np.float32(1.0))
        return torch.tensor([tot], dtype=torch.int64)
    def generate(self, total, offset, r, m_lo, count):
        idx = oracle.systematic(self.q, total, offset, r, self.n_global, m_lo, count)
        assert (idx >= 0).all()
        out = self.rows[idx].copy(); out[:, 3] = 1.0 / self.n_global
        return torch.from_numpy(out)
    def inbox(self, rows):
        self._in = torch.empty(rows, 4, dtype=torch.float32); return self._in
    def adopt(self, rows): self.rows = self._in.numpy().copy()

class OraclePullShard(OracleShard):
    """Pull mode: peer memory is stood in for by an all-gather of every rank's CDF increments and rows
    (what the kernel would read through NVLink pointers); the collectives and their order are the real ones."""
    def connect_peers(self, rank, world, n_global, group=None): self.rank, self.world = rank, world
    def totals_all(self, world):
        self._tot = torch.zeros(world, dtype=torch.int64); return self._tot
    def pull(self, u0):
        qs, rows = [None] * self.world, [None] * self.world
        dist.all_gather_object(qs, self.q); dist.all_gather_object(rows, self.rows)
        q, allrows = np.concatenate(qs), np.concatenate(rows)
        total = int(self._tot.numpy().view(np.uint64).sum(dtype=np.uint64))
        assert total == int(q.sum(dtype=np.uint64))
        lo, cnt = slots_of(self.rank, self.world, self.n_global)
        idx = oracle.systematic(q, total, 0, oracle.systematic_offset(u0, total), self.n_global, lo, cnt)
        self.rows = allrows[idx].copy(); self.rows[:, 3] = 1.0 / self.n_global

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = int(os.environ["AMCLJ_N"])
rng = np.random.Generator(np.random.PCG64(42))
rows = rng.random((n, 4)).astype(np.float32)
rows[:, 0] = np.arange(n)                      # x carries the global index
raw = (np.log(rng.random(n) ** 8 + 1e-30)).astype(np.float32)
if os.environ.get("AMCLJ_SKEW") == "1":
    raw[: n // 2] -= 60.0                      # nearly all weight on the last rank
lo, cnt = slots_of(rank, world, n)
mode = os.environ.get("AMCLJ_MODE", "a2a")
be = (OraclePullShard if mode == "pull" else OracleShard)(rows[lo:lo + cnt], raw[lo:lo + cnt])
f = ShardedFilter(None, rank, world, n, None, backend=be, mode=mode)
u0 = 0.4142135
f.update((0, 0, 0), (0, 0, 0), 1, u0)
got = [None] * world
dist.all_gather_object(got, be.rows[:, 0].astype(np.int64))
if rank == 0:
    lin, wmax = oracle.linear_weights(raw, True)
    q, total = oracle.fixed_weights(lin, wmax)
    ref = oracle.systematic(q, total, 0, oracle.systematic_offset(u0, total), n, 0, n)
    assert np.array_equal(np.concatenate(got), ref), "sharded index list differs from the single list"
    print("SHARDED_OK", n, world)
dist.destroy_process_group()
'''


def _run(n, skew, port, mode="a2a"):
    env = dict(os.environ, AMCLJ_ROOT=str(ROOT), AMCLJ_N=str(n), AMCLJ_SKEW="1" if skew else "0", AMCLJ_MODE=mode)
    script = ROOT / "tests" / f"_sharded_worker_{port}.py"
    script.write_text(WORKER)
    try:
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                           env=env, capture_output=True, text=True, timeout=300)
    finally:
        script.unlink(missing_ok=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "SHARDED_OK" in p.stdout


def test_sharded_update_world2_balanced():
    _run(5001, False, 29731)


def test_sharded_update_world2_skewed():
    _run(4000, True, 29732)


def test_sharded_update_world2_pull_choreography():
    """the pull choreography (max all-reduce, totals all-gather, then every rank fills its own
    slots from whichever rank holds them) on CPU, balanced and skewed."""
    _run(5001, False, 29733, mode="pull")
    _run(4000, True, 29734, mode="pull")
