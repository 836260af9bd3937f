"""world_size-2 gloo test of the N>1 host path: weight broadcast + batch sharding (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from s2m2_b200.parallel import shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from s2m2_b200.model import S2M2
    from s2m2_b200.parallel import broadcast_weights, flat_weights
    from s2m2_b200.weights import random_state_dict, synthetic_pair
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from torch_ops import TorchOps
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = S2M2(32, 1, 1, use_positivity=True, refine_iter=1, precision="fp32").eval()
    if rank == 0:
        m.my_load_state_dict(random_state_dict(32, 1, 0))
    nbytes = broadcast_weights(m, 0)
    ref = torch.cat([v.reshape(-1) for v in random_state_dict(32, 1, 0).values()])
    same_w = bool(torch.equal(flat_weights(m), ref))
    # every rank runs the same pair through the same schedule -> bit-identical outputs (rank consistency)
    m.ops = TorchOps()
    l, r = synthetic_pair(1, 64, 64, 3)
    d, o, c = m(l, r)
    gathered = [torch.zeros_like(d) for _ in range(world)]
    dist.all_gather(gathered, d)
    q.put((rank, nbytes, same_w, bool(torch.equal(gathered[0], gathered[1]))))
    dist.destroy_process_group()


def _worker_packed(rank, world, port, q):
    """Rank 0 owns the weights; ONE broadcast of the PACKED blob; rank 1 never sees the fp32 state-dict."""
    import torch.distributed as dist
    from s2m2_b200.model import S2M2
    from s2m2_b200.parallel import broadcast_packed
    from s2m2_b200.weights import random_state_dict, synthetic_pair
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from torch_ops import TorchOps
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = S2M2(32, 1, 1, use_positivity=True, refine_iter=2, precision="fp16").eval()
    if rank == 0:
        m.my_load_state_dict(random_state_dict(32, 1, 0))
    m.ops = TorchOps()
    nbytes = broadcast_packed(m, torch.device("cpu"), 0)
    # the fp32 conv / linear weights never travel (norm weights and biases are fp32 in the packed form too and alias the parameters)
    zero_params = float(m._p["cnn_backbone/conv2/0/weight"].abs().max()) == 0.0 and float(m._p["ctx_feat/0/weight"].abs().max()) == 0.0
    l, r = synthetic_pair(1, 64, 64, 3)
    d, o, c = m(l, r)
    gathered = [torch.zeros_like(d) for _ in range(world)]
    dist.all_gather(gathered, d)
    # single-process result with the weights loaded the ordinary way
    m1 = S2M2(32, 1, 1, use_positivity=True, refine_iter=2, precision="fp16").eval()
    m1.my_load_state_dict(random_state_dict(32, 1, 0))
    m1.ops = TorchOps()
    d1, _, _ = m1(l, r)
    q.put((rank, nbytes, zero_params, bool(torch.equal(gathered[0], gathered[1])), bool(torch.equal(d, d1))))
    dist.destroy_process_group()


def test_packed_blob_broadcast_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nbytes, zero_params, same_out, same_as_single in res:
        assert nbytes > 500_000 and same_out and same_as_single, (rank, nbytes, same_out, same_as_single)
        assert zero_params == (rank != 0)


def test_weight_broadcast_and_rank_consistency_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nbytes, same_w, same_out in res:
        assert nbytes > 1_000_000 and same_w and same_out, (rank, nbytes, same_w, same_out)
