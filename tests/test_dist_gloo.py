"""world_size-2 gloo test of the multi-GPU host logic: utterances are sharded by
length, each rank runs the path on its shard alone (here: the CPU oracle stands in for the rank's
GPU), rank results are all_gathered, and the union must equal the single-process result bit for
bit — there is no collective on the data path, so sharding may not change any value."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from telugu_asr_b200 import shard_by_length


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lens, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wav, ln = oracle.make_waveforms(lens, seed=21, dist="tilt")
        mine = shard_by_length(ln, world)[rank]
        weights = oracle.glorot_subsampling_weights(32, 80, seed=3)
        feat, nf = oracle.logmel_batch_ref(wav[mine][:, : int(ln[mine].max())], ln[mine])
        out, mask, len_all = oracle.subsample_ref(feat, nf, weights)
        # per-utterance checksum over VALID positions + lengths, gathered for validation only
        rows = []
        for j, i in enumerate(mine):
            L3 = max(int(len_all[-1][j]), 0)
            rows.append([float(i), float(nf[j]), float(len_all[-1][j]), float(out[j, :L3].astype(np.float64).sum())])
        t = torch.tensor(rows, dtype=torch.float64).reshape(-1, 4)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64))
        mx = int(max(s.item() for s in sizes))
        pad = torch.zeros((mx, 4), dtype=torch.float64)
        pad[: t.shape[0]] = t
        bufs = [torch.zeros((mx, 4), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(bufs, pad)
        dist.barrier()
        if rank == 0:
            allrows = torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)]).numpy()
            q.put(allrows)
    finally:
        dist.destroy_process_group()


def test_shard_union_equals_single_process():
    lens = [16000, 9000, 30000, 4000, 22000, 12000, 399]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = rows[np.argsort(rows[:, 0])]
    assert rows[:, 0].astype(int).tolist() == list(range(len(lens)))
    # single-process reference on the whole batch
    wav, ln = oracle.make_waveforms(lens, seed=21, dist="tilt")
    weights = oracle.glorot_subsampling_weights(32, 80, seed=3)
    feat, nf = oracle.logmel_batch_ref(wav, ln)
    out, mask, len_all = oracle.subsample_ref(feat, nf, weights)
    np.testing.assert_array_equal(rows[:, 1].astype(np.int64), nf)
    np.testing.assert_array_equal(rows[:, 2].astype(np.int64), len_all[-1])
    for i in range(len(lens)):
        L3 = max(int(len_all[-1][i]), 0)
        assert rows[i, 3] == float(out[i, :L3].astype(np.float64).sum())
