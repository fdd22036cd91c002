"""N > 1 host logic on CPU: world-size-2 gloo processes own contiguous batch shards (dldt_b200.multi.shard_range), run their
shard, and the gathered result - including the re-packed DetectionOutput tensor - must equal the unsharded run.  The device
is replaced by the numpy oracle here (no GPU in this container); the scatter/gather code is the product's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dldt_b200 import multi, zoo
        from oracle import interp
        N = 5 if case == "resnet" else 4
        rng = np.random.default_rng(99)
        x = rng.uniform(0, 1, (N, 3, 64, 64) if case == "resnet" else (N, 3, 96, 96)).astype(np.float32)

        def build(b):
            if case == "resnet":
                return zoo.resnet50(batch=b, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
            return zoo.ssd300_resnet34(batch=b, image=96, width=16, num_classes=6, trunk_blocks=(1, 1, 1), keep_top_k=20, top_k=10)

        b, e = multi.shard_range(N, world, rank)
        out = interp.run(build(e - b), {"data": x[b:e]})
        name = next(iter(out))
        gathered = [None] * world
        dist.all_gather_object(gathered, (b, out[name]))
        # max-over-ranks reduction used by bench.py for its timings
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            merged = multi.merge_outputs(name, [g[1] for g in gathered], [g[0] for g in gathered])
            full = interp.run(build(N), {"data": x})[name]
            q.put((merged, full, float(t.item()), [g[0] for g in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["resnet", "ssd"])
def test_two_rank_batch_sharding_matches_unsharded(case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, full, tmax, offsets = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    assert merged.shape == full.shape
    if case == "ssd":
        m, f = merged.reshape(-1, 7), full.reshape(-1, 7)
        np.testing.assert_array_equal(m[:, :2], f[:, :2])      # image ids (renumbered), labels, -1 terminator
        np.testing.assert_allclose(m[:, 2:], f[:, 2:], rtol=1e-5, atol=1e-6)
    else:
        assert offsets == [0, 3]
        np.testing.assert_allclose(merged, full, rtol=1e-4, atol=1e-5)


def test_shard_range_covers_the_batch_exactly():
    from dldt_b200.multi import shard_range
    for n in (1, 7, 8, 32, 255, 256):
        for w in (1, 2, 3, 4, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_merge_detection_outputs_edge_cases():
    from dldt_b200.multi import merge_detection_outputs
    a = np.zeros((1, 1, 4, 7), np.float32)
    a[0, 0, 0] = [0, 3, .9, 0, 0, 1, 1]
    a[0, 0, 1] = [1, 2, .8, 0, 0, 1, 1]
    a[0, 0, 2, 0] = -1
    b = np.zeros((1, 1, 4, 7), np.float32)
    b[0, 0, 0, 0] = -1                                          # shard with no detections at all
    c = np.zeros((1, 1, 2, 7), np.float32)
    c[0, 0, 0] = [0, 1, .7, 0, 0, 1, 1]
    c[0, 0, 1] = [0, 5, .6, 0, 0, 1, 1]                         # completely full shard: no terminator row
    m = merge_detection_outputs([a, b, c], [0, 2, 4]).reshape(-1, 7)
    assert m.shape == (10, 7)
    assert m[:, 0].tolist() == [0, 1, 4, 4, -1, 0, 0, 0, 0, 0]
    assert m[:4, 1].tolist() == [3, 2, 1, 5]
