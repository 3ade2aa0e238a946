"""world_size-2 gloo test of the multi-GPU path's host logic: contiguous batch shards, descriptor
broadcast, max-over-ranks timing reduction and the gather of results.  The per-shard compute is stood in by
the CPU oracle (this is a test: the GPU product path is exercised by tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, n, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import wavelets_b200 as wb
        from wavelets_b200.shard import shard_range, broadcast_descriptor, allreduce_max, allreduce_sum, gather_columns, scatter_columns
        from oracle import oracle as orc
        qmf = wb.wavelet(wb.WT.db4).qmf if rank == 0 else None
        qmf = broadcast_descriptor(qmf, src=0)                      # only rank 0 built the descriptor
        x = np.random.default_rng(42).standard_normal((n, B))       # same synthetic batch on every rank
        lo, hi = shard_range(B, rank, world)
        y_local = orc.dwt_filter_batch(x[:, lo:hi], 1, qmf, 6)      # this rank's contiguous block of columns
        t_rank = 1.0 + rank                                         # pretend device time
        t_max = allreduce_max(t_rank)
        checksum = allreduce_sum(float(np.sum(y_local)))
        full = gather_columns(torch.tensor(np.ascontiguousarray(y_local)), B, dst=0)
        # the batch starting on rank 0 only: scatter (batch-major storage = column-major (n, B)), transform, gather
        xb = torch.tensor(np.ascontiguousarray(x.T)) if rank == 0 else None
        mine = scatter_columns(xb, B, (n,), torch.float64, "cpu", src=0)
        y2 = orc.dwt_filter_batch(np.ascontiguousarray(mine.numpy().T), 1, qmf, 6)
        full2 = gather_columns(torch.tensor(np.ascontiguousarray(y2)), B, dst=0)
        if rank == 0:
            ref = orc.dwt_filter_batch(x, 1, qmf, 6)
            ok = (t_max == float(world)) and np.array_equal(full.numpy(), ref) and abs(checksum - ref.sum()) < 1e-9
            ok = ok and np.array_equal(full2.numpy(), ref)
            open(out_path, "w").write("ok" if ok else "bad")
    finally:
        dist.destroy_process_group()


def test_two_rank_batch_split(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), 7, 64, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
