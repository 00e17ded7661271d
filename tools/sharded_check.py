"""torchrun worker: sharded filter update on G GPUs vs the same update on one GPU.

  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/sharded_check.py [N] [mailbox|pull|a2a] [spread|dense]

`dense` shards a pose-tracking cloud (C2's gaussian), which every rank runs on the per-cell ray
tables (three updates, so that the policy's hint has settled); `spread` (default) is C3's uniform cloud.

Every rank runs its shard through amclj_b200.sharded.ShardedFilter (mailbox: one library call per
update, exchanges inside the kernels; pull / a2a: NCCL all-reduce(max), all-gather(totals), then peer pull
or all-to-all rebalance); rank 0 also runs all N particles on its own GPU
through amclj_filter_update and compares: identical particle lists (same source index for
every output slot) and identical motion-sampled poses (Philox keyed by global particle id).
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amclj_b200 import _lib, workloads  # noqa: E402
from amclj_b200.sharded import ShardedFilter, slots_of  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    mode = sys.argv[2] if len(sys.argv) > 2 else "mailbox"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cloud = sys.argv[3] if len(sys.argv) > 3 else "spread"
    w = workloads.c2(n) if cloud == "dense" else workloads.c3(n)
    updates = 3 if cloud == "dense" else 2
    s = w.scan
    lo, cnt = slots_of(rank, world, n)
    ctx = _lib.Context(local, "NS")
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_map(w.grid)
    ctx.set_shard(rank, world, lo, n)
    ctx.set_particles(w.x[lo:lo + cnt], w.y[lo:lo + cnt], w.theta[lo:lo + cnt])
    with torch.cuda.stream(stream):
        f = ShardedFilter(ctx, rank, world, n, dev, mode=mode)
        f.stage_scan(s)
        for it in range(updates):  # consecutive updates: the later ones start from rebalanced shards
            f.update(w.curr, w.prev, seed=77 + it, u0=w.u0)
        torch.cuda.synchronize()
    x, y, t, wt = ctx.get_particles()
    rows = np.stack([x, y, t, wt], axis=1)
    got = [None] * world
    dist.all_gather_object(got, rows)
    ok = True
    if rank == 0:
        ref = _lib.Context(local, "NS")
        ref.set_map(w.grid)
        ref.set_particles(w.x, w.y, w.theta)
        for it in range(updates):
            ref.filter_update(w.curr, w.prev, s.ranges, s.angle_min, s.angle_inc, s.range_max, w.u0, seed=77 + it)
        rx, ry, rt, rw = ref.get_particles()
        full = np.concatenate(got, axis=0)
        ok = (np.array_equal(full[:, 0], rx) and np.array_equal(full[:, 1], ry) and np.array_equal(full[:, 2], rt)
              and np.allclose(full[:, 3], rw))
        print(f"SHARDED_GPU_{'OK' if ok else 'MISMATCH'} mode={mode} cloud={cloud} world={world} n={n} "
              f"mismatching_slots={(full[:, 0] != rx).sum()}", flush=True)
        ref.close()
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
