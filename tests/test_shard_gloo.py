"""Multi-rank host logic on CPU: world_size 2 over gloo. The data path has no collective (images are
independent); what is exercised here is the shard assignment and the gather of blob sizes, with the ORACLE
standing in for the GPU encoder (this is a test: the product path never calls the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ako_b200 import shard
import oracle_lib as ol


def test_shard_indices_cover_and_balance():
    for n in (0, 1, 7, 8, 4096):
        for world in (1, 2, 4, 8):
            parts = [shard.shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
            assert all(shard.owner_of(i, world) == r for r, p in enumerate(parts) for i in p)
    assert shard.blob_offsets([5, 3, 9], align=4).tolist() == [0, 8, 12]


def _worker(rank, world, port, n_images, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = ol.load_oracle()
    images = [ol.synth(orc, 48, 40, 1000 + i) for i in range(n_images)]

    def encode_one(img):
        return ol.orc_encode(orc, img, wavelet=1, q=8, g=0)[0]

    blobs, sizes = shard.encode_sharded(encode_one, images, rank, world, dist)
    # every rank knows every size; the blobs of the batch can be laid out without moving pixel data
    want = [len(encode_one(im)) for im in images]
    ok = sizes.tolist() == want and sorted(blobs) == shard.shard_indices(n_images, rank, world)
    ok = ok and all(len(blobs[i]) == want[i] for i in blobs)
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out[rank] = int(t.item())
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ol.build_oracle()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, 5, out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}
