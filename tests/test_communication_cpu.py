"""The 20 blocking collectives + process groups on gloo (world 3) against torch.distributed / closed forms — the strategy
of the reference's tests/comm/test_communicator.py and examples/communication_primitives/main.py."""
import torch

from tests.mp_utils import run_distributed


def _primitives(rank, world):
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200 import ReduceOp

    bagua.init_process_group()
    assert bagua.is_initialized() and bagua.get_rank() == rank and bagua.get_world_size() == world
    comm = bagua.communication._get_default_group().get_global_communicator()
    assert comm.rank() == rank and comm.nranks() == world and comm.device_id() == -1 and not comm.check_abort()
    base = torch.arange(6, dtype=torch.float32)

    t = base * (rank + 1)
    bagua.allreduce_inplace(t)
    assert torch.equal(t, base * sum(range(1, world + 1)))
    t = base * (rank + 1)
    out = torch.zeros(6)
    bagua.allreduce(t, out, op=ReduceOp.AVG)
    torch.testing.assert_close(out, base * (world + 1) / 2)
    for op, fn in [(ReduceOp.MAX, max), (ReduceOp.MIN, min)]:
        t = torch.full((3,), float(rank))
        bagua.allreduce_inplace(t, op=op)
        assert t[0].item() == fn(range(world))
    t = torch.full((2,), 2.0)
    bagua.allreduce_inplace(t, op=ReduceOp.PRODUCT)
    assert t[0].item() == 2.0 ** world
    ti = torch.tensor([1 << rank])
    bagua.allreduce_inplace(ti, op=ReduceOp.BOR)
    assert ti.item() == (1 << world) - 1
    ts = [torch.ones(3) * rank, torch.ones(2, dtype=torch.int64) * rank]
    bagua.allreduce_coalesced_inplace(ts)
    assert ts[0][0].item() == sum(range(world)) and ts[1][0].item() == sum(range(world))

    t = torch.full((4,), float(rank))
    bagua.broadcast(t, src=1)
    assert t[0].item() == 1.0
    obj = bagua.broadcast_object({"a": rank} if rank == 2 else None, src=2)
    assert obj == {"a": 2}
    a, b = torch.full((2,), float(rank)), torch.full((3,), float(rank) + 10)
    bagua.broadcast_coalesced([a, b], src=0)
    assert a[0].item() == 0.0 and b[0].item() == 10.0

    send, recv = torch.full((3,), float(rank + 1)), torch.zeros(3)
    bagua.reduce(send, recv, dst=0)
    if rank == 0:
        assert recv[0].item() == sum(range(1, world + 1))
    t = torch.full((3,), float(rank + 1))
    bagua.reduce_inplace(t, dst=1, op=ReduceOp.AVG)
    if rank == 1:
        assert abs(t[0].item() - (world + 1) / 2) < 1e-6

    send, recv = torch.full((2,), float(rank)), torch.zeros(2 * world)
    bagua.allgather(send, recv)
    assert recv.tolist() == [float(r) for r in range(world) for _ in range(2)]
    t = torch.zeros(2 * world)
    t[2 * rank : 2 * rank + 2] = rank + 5
    bagua.allgather_inplace(t)
    assert t.tolist() == [float(r + 5) for r in range(world) for _ in range(2)]

    send, recv = torch.full((2,), float(rank)), torch.zeros(2 * world)
    bagua.gather(send, recv, dst=2)
    if rank == 2:
        assert recv.tolist() == [float(r) for r in range(world) for _ in range(2)]
    src_t = torch.arange(2 * world, dtype=torch.float32)
    recv = torch.zeros(2)
    bagua.scatter(src_t, recv, src=0)
    assert recv.tolist() == [2.0 * rank, 2.0 * rank + 1]

    send, recv = torch.arange(2 * world, dtype=torch.float32) * (rank + 1), torch.zeros(2)
    bagua.reduce_scatter(send, recv)
    s = sum(range(1, world + 1))
    assert recv.tolist() == [2.0 * rank * s, (2.0 * rank + 1) * s]
    t = torch.arange(2 * world, dtype=torch.float32) * (rank + 1)
    bagua.reduce_scatter_inplace(t)
    assert t[:2].tolist() == [2.0 * rank * s, (2.0 * rank + 1) * s]

    send, recv = torch.tensor([float(rank * 10 + p) for p in range(world)]), torch.zeros(world)
    bagua.alltoall(send, recv)
    assert recv.tolist() == [float(p * 10 + rank) for p in range(world)]
    t = torch.tensor([float(rank * 10 + p) for p in range(world)])
    bagua.alltoall_inplace(t)
    assert t.tolist() == [float(p * 10 + rank) for p in range(world)]
    # alltoall_v: rank r sends (p+1) elements to peer p
    counts = [p + 1 for p in range(world)]
    displs = [sum(counts[:p]) for p in range(world)]
    send = torch.cat([torch.full((p + 1,), float(rank * 100 + p)) for p in range(world)])
    rcounts = [rank + 1] * world
    rdispls = [p * (rank + 1) for p in range(world)]
    recv = torch.zeros(world * (rank + 1))
    bagua.alltoall_v(send, counts, displs, recv, rcounts, rdispls)
    assert recv.tolist() == [float(p * 100 + rank) for p in range(world) for _ in range(rank + 1)]

    if rank == 0:
        bagua.send(torch.tensor([42.0]), dst=1)
    elif rank == 1:
        r = torch.zeros(1)
        bagua.recv(r, src=0)
        assert r.item() == 42.0
    bagua.barrier()

    # sub-groups and torch-group wrapping
    sub = bagua.new_group(ranks=[0, 2])
    if rank in (0, 2):
        t = torch.tensor([float(rank)])
        bagua.allreduce_inplace(t, comm=sub.get_global_communicator())
        assert t.item() == 2.0
        assert sub.get_global_communicator().nranks() == 2
    tg = dist.new_group([0, 1, 2])
    bg = bagua.from_torch_group(tg)
    assert bagua.from_torch_group(tg) is bg and tg.bagua_pg is bg
    assert tg.bagua_get_global_communicator().nranks() == 3
    assert bg.get_intra_node_communicator().nranks() == 3 and bg.get_inter_node_communicator().nranks() == 1
    # autograd-aware allreduce
    from bagua_b200.parallel.data_parallel import functional

    x = torch.ones(2, requires_grad=True)
    y = functional.all_reduce(x * (rank + 1))
    y.sum().backward()
    assert y[0].item() == sum(range(1, world + 1)) and x.grad[0].item() == world * (rank + 1)
    return True


def test_blocking_collectives_and_groups():
    assert all(run_distributed(_primitives, world=3, timeout=300))


def test_init_twice_raises_and_uninitialised_default_group():
    import pytest

    import bagua_b200 as bagua

    assert not bagua.is_initialized()
    with pytest.raises(RuntimeError):
        bagua.communication._get_default_group()


def _virtual_nodes_worker(rank, world):
    """4 processes pretending to be 2 nodes x 2 ranks: intra / rail groups and the hierarchical all-reduce legs."""
    import os

    os.environ["NODE_RANK"] = str(rank // 2)
    os.environ["LOCAL_RANK"] = str(rank % 2)
    os.environ["LOCAL_WORLD_SIZE"] = "2"
    import torch

    import bagua_b200 as bagua
    from bagua_b200.bucket import _torch_allreduce

    bagua.init_process_group()
    pg = bagua.communication._get_default_group()
    assert pg.nnodes == 2 and pg.intra_ranks == [2 * (rank // 2), 2 * (rank // 2) + 1] and pg.inter_ranks == [rank % 2, rank % 2 + 2]
    assert pg.peer_engine() is None and pg.hier_engine() is None           # CPU: no peer kernels, torch legs instead
    t = torch.full((10,), float(rank + 1))
    _torch_allreduce(t, pg, True, True)                                    # intra reduce → inter all-reduce → intra broadcast
    sub = bagua.new_group(ranks=[0, 3])                                    # one rank per node, different local ranks: rails must still connect them
    if rank in (0, 3):
        assert sub.nnodes == 2 and sub.intra_ranks == [rank] and sub.inter_ranks == [0, 3]
        u = torch.full((4,), float(rank + 1))
        _torch_allreduce(u, sub, True, True)
        assert torch.allclose(u, torch.full((4,), 2.5))
    return t


def test_virtual_multi_node_groups_and_hierarchical_allreduce():
    for t in run_distributed(_virtual_nodes_worker, world=4):
        assert torch.allclose(t, torch.full((10,), 2.5))


def _explicit_store_worker(rank, world):
    """``init_process_group(store=..., rank=..., world_size=..., local_world_size=...)``: no env:// rendezvous involved."""
    import datetime
    import os

    import torch
    import torch.distributed as dist

    import bagua_b200 as bagua

    port = int(os.environ["MASTER_PORT"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):   # nothing may be read from the launcher environment
        os.environ.pop(k, None)
    store = dist.TCPStore("127.0.0.1", port, world, is_master=(rank == 0), timeout=datetime.timedelta(seconds=60))
    bagua.init_process_group(store=store, rank=rank, world_size=world, local_world_size=world)
    assert bagua.get_rank() == rank and bagua.get_world_size() == world and bagua.get_local_size() == world
    t = torch.full((3,), float(rank))
    bagua.allreduce_inplace(t, op=bagua.ReduceOp.SUM)
    objs = bagua.broadcast_object({"from": rank}, src=1)
    return t, objs


def test_init_process_group_with_an_explicit_store():
    res = run_distributed(_explicit_store_worker, world=2)
    for t, obj in res:
        assert torch.equal(t, torch.full((3,), 1.0)) and obj == {"from": 1}
