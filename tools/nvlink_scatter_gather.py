#!/usr/bin/env python3
"""SURVEY 8e, optional row: the batch starts (and ends) on ONE GPU.  Rank 0 scatters contiguous blocks of columns to its
peers with NCCL send/recv (peer copies over NVLink / NVSwitch), every rank runs the single-GPU path on its shard, rank 0
gathers the results and compares them bit for bit with its own single-GPU transform of the whole batch.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nvlink_scatter_gather.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import wavelets_b200 as wb
from wavelets_b200.shard import scatter_columns, shard_sizes

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, B = 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wt = wb.wavelet(wb.WT.db4)
full = torch.randn((B, n), device=dev) if rank == 0 else None           # storage of the column-major (n, B) batch
for it in range(3):
    dist.barrier(); torch.cuda.synchronize(dev); t0 = time.perf_counter()
    mine = scatter_columns(full, B, (n,), torch.float32, dev, src=0)
    torch.cuda.synchronize(dev); dist.barrier(); t1 = time.perf_counter()
    y = wb.dwtc(mine.t(), wt)                                            # (n, b_r) column-major view of the shard
    torch.cuda.synchronize(dev); dist.barrier(); t2 = time.perf_counter()
    # gather: every shard's storage is one contiguous block of the (n, B) column-major result -> received in place
    ys = y.t()                                                           # (b_r, n) contiguous = the shard's storage
    if rank == 0:
        outb = torch.empty((B, n), device=dev)
        lo = 0
        for r, sz in enumerate(shard_sizes(B, world)):
            if r == 0:
                outb[lo:lo + sz].copy_(ys)
            elif sz:
                dist.recv(outb[lo:lo + sz], src=r)
            lo += sz
        out = outb.t()
    elif ys.shape[0]:
        dist.send(ys, dst=0)
    torch.cuda.synchronize(dev); dist.barrier(); t3 = time.perf_counter()
if rank == 0:
    ref = wb.dwtc(full.t(), wt)
    moved = 4.0 * n * (B - shard_sizes(B, world)[0])
    print(json.dumps({"n_gpus": world, "columns": B, "identical_to_single_gpu": bool(torch.equal(out, ref)),
                      "scatter_ms": (t1 - t0) * 1e3, "transform_ms": (t2 - t1) * 1e3, "gather_ms": (t3 - t2) * 1e3,
                      "scatter_gbs": moved / (t1 - t0) / 1e9, "gather_gbs": moved / (t3 - t2) / 1e9,
                      "note": "NCCL send/recv between the GPUs of one box (NVLink 5 / NVSwitch peer copies); bytes = the columns that leave rank 0"}))
dist.destroy_process_group()
