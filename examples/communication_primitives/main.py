"""Every blocking communication primitive of the public API, each checked against ``torch.distributed`` where torch has the
same collective and against a closed-form expectation everywhere (the reference's examples/communication_primitives/main.py:14-182
walks send/recv, broadcast, allreduce, reduce, allgather, reduce_scatter, alltoall, alltoall_v; this one covers the whole list of
bagua/torch_api/communication.py:573-1401, including the in-place and coalesced flavours, ``broadcast_object`` and sub-groups).

    python -m bagua_b200.distributed.launch --nproc_per_node=8 examples/communication_primitives/main.py

Rank ``r`` contributes ``data(r) = arange(L)·(r + 1) + r``, so every rank can compute every expected result locally.  CUDA tensors
on one NVSwitch node take the peer kernels for all-reduce; everything else (and the gloo run on CPUs) goes through
torch.distributed on the group's communication stream."""
import logging
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bagua_b200 as bagua  # noqa: E402
from bagua_b200 import ReduceOp  # noqa: E402


def main():
    cuda = torch.cuda.is_available() and os.environ.get("BAGUA_FORCE_CPU", "0") != "1"
    if cuda:
        torch.cuda.set_device(bagua.get_local_rank())
    bagua.init_process_group()
    rank, n = bagua.get_rank(), bagua.get_world_size()
    assert n >= 2, "world size must be at least 2"
    logging.basicConfig(format="%(message)s", level=logging.INFO if rank == 0 else logging.ERROR, stream=sys.stdout)
    dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
    comm = bagua.communication._get_default_group().get_global_communicator()
    L = 4 * n   # divisible by the world size: the chunked collectives split it evenly

    def data(r, length=L):
        return torch.arange(length, dtype=torch.float32, device=dev) * (r + 1) + r

    def same(got, want, what):
        assert torch.allclose(got, want), f"{what}: rank {rank} got {got.tolist()} expected {want.tolist()}"
        logging.info("ok  %s", what)

    total = sum(data(r) for r in range(n))
    chunk = L // n
    mine = slice(rank * chunk, (rank + 1) * chunk)

    # ---- point to point (ring: every rank sends to its right neighbour; even ranks send first)
    got = torch.zeros(L, device=dev)
    right, left = (rank + 1) % n, (rank - 1) % n
    if rank % 2 == 0:
        bagua.send(data(rank), right, comm=comm)
        bagua.recv(got, left, comm=comm)
    else:
        bagua.recv(got, left, comm=comm)
        bagua.send(data(rank), right, comm=comm)
    same(got, data(left), "send / recv")

    # ---- broadcast family
    t = data(rank)
    bagua.broadcast(t, 1, comm=comm)
    ref = data(rank)
    dist.broadcast(ref, 1)
    same(t, ref, "broadcast vs torch.distributed")
    same(t, data(1), "broadcast")
    ts = [data(rank, 3), data(rank, 5).double(), data(rank, 2)]
    bagua.broadcast_coalesced(ts, 0, comm=comm)
    same(torch.cat([x.float() for x in ts]), torch.cat([data(0, 3), data(0, 5), data(0, 2)]), "broadcast_coalesced (mixed dtypes)")
    obj = bagua.broadcast_object({"from": rank, "payload": list(range(rank + 3))} if rank == n - 1 else None, n - 1, comm=comm)
    assert obj == {"from": n - 1, "payload": list(range(n + 2))}, obj
    logging.info("ok  broadcast_object")

    # ---- reductions
    out = torch.zeros(L, device=dev)
    bagua.allreduce(data(rank), out, comm=comm)
    ref = data(rank)
    dist.all_reduce(ref)
    same(out, ref, "allreduce vs torch.distributed")
    same(out, total, "allreduce")
    t = data(rank)
    bagua.allreduce_inplace(t, op=ReduceOp.AVG, comm=comm)
    same(t, total / n, "allreduce_inplace (AVG)")
    t = data(rank)
    bagua.allreduce_inplace(t, op=ReduceOp.MAX, comm=comm)
    same(t, data(n - 1), "allreduce_inplace (MAX)")
    ts = [data(rank, 3), data(rank, 6).double()]
    bagua.allreduce_coalesced_inplace(ts, comm=comm)
    same(ts[1].float(), sum(data(r, 6) for r in range(n)), "allreduce_coalesced_inplace")
    out = torch.full((L,), -1.0, device=dev)
    bagua.reduce(data(rank), out, dst=1, comm=comm)
    same(out, total if rank == 1 else torch.full((L,), -1.0, device=dev), "reduce")
    t = data(rank)
    bagua.reduce_inplace(t, dst=0, op=ReduceOp.MIN, comm=comm)
    if rank == 0:
        same(t, data(0), "reduce_inplace (MIN)")
    out = torch.zeros(chunk, device=dev)
    bagua.reduce_scatter(data(rank), out, comm=comm)
    same(out, total[mine], "reduce_scatter")
    t = data(rank)
    bagua.reduce_scatter_inplace(t, comm=comm)
    same(t[:chunk], total[mine], "reduce_scatter_inplace (result in the first chunk)")

    # ---- gather / scatter family
    everyone = torch.cat([data(r, chunk) for r in range(n)])
    out = torch.zeros(L, device=dev)
    bagua.allgather(data(rank, chunk), out, comm=comm)
    pieces = [torch.zeros(chunk, device=dev) for _ in range(n)]
    dist.all_gather(pieces, data(rank, chunk))
    same(out, torch.cat(pieces), "allgather vs torch.distributed")
    same(out, everyone, "allgather")
    t = torch.zeros(L, device=dev)
    t[mine] = data(rank, chunk)
    bagua.allgather_inplace(t, comm=comm)
    same(t, everyone, "allgather_inplace")
    out = torch.full((L,), -1.0, device=dev)
    bagua.gather(data(rank, chunk), out, dst=1, comm=comm)
    same(out, everyone if rank == 1 else torch.full((L,), -1.0, device=dev), "gather")
    t = torch.zeros(L, device=dev)
    t[:chunk] = data(rank, chunk)
    bagua.gather_inplace(t, chunk, dst=0, comm=comm)
    if rank == 0:
        same(t, everyone, "gather_inplace")
    out = torch.zeros(chunk, device=dev)
    bagua.scatter(data(7), out, src=0, comm=comm)   # only src's send buffer matters
    same(out, data(7)[mine], "scatter")
    t = data(7) if rank == 1 else torch.zeros(L, device=dev)
    bagua.scatter_inplace(t, chunk, src=1, comm=comm)
    same(t[:chunk], data(7)[mine], "scatter_inplace")

    # ---- all-to-all family: element j of chunk i on rank r is 100·r + 10·i + j
    def a2a_send(r):
        return (100.0 * r + 10.0 * torch.arange(n, device=dev).view(n, 1) + torch.arange(chunk, device=dev).view(1, chunk)).reshape(-1)

    want = torch.cat([a2a_send(src).view(n, chunk)[rank] for src in range(n)])
    out = torch.zeros(L, device=dev)
    bagua.alltoall(a2a_send(rank), out, comm=comm)
    same(out, want, "alltoall")
    t = a2a_send(rank)
    bagua.alltoall_inplace(t, comm=comm)
    same(t, want, "alltoall_inplace")
    # variable counts: rank r sends (d + 1) elements to rank d, so rank d receives (d + 1) elements from everybody
    send_counts = [d + 1 for d in range(n)]
    send_displs = [sum(send_counts[:d]) for d in range(n)]
    recv_counts = [rank + 1] * n
    recv_displs = [i * (rank + 1) for i in range(n)]
    sendbuf = torch.cat([torch.full((d + 1,), float(100 * rank + d), device=dev) for d in range(n)])
    out = torch.zeros(n * (rank + 1), device=dev)
    bagua.alltoall_v(sendbuf, send_counts, send_displs, out, recv_counts, recv_displs, comm=comm)
    same(out, torch.cat([torch.full((rank + 1,), float(100 * src + rank), device=dev) for src in range(n)]), "alltoall_v")
    counts, displs = [2] * n, [2 * i for i in range(n)]
    t = torch.cat([torch.full((2,), float(100 * rank + d), device=dev) for d in range(n)])
    bagua.alltoall_v_inplace(t, counts, displs, comm=comm)
    same(t, torch.cat([torch.full((2,), float(100 * src + rank), device=dev) for src in range(n)]), "alltoall_v_inplace")

    # ---- a sub-group has its own communicators (and its own AVG divisor)
    evens = bagua.communication.new_group(ranks=list(range(0, n, 2)))
    if rank % 2 == 0:
        t = data(rank)
        bagua.allreduce_inplace(t, op=ReduceOp.AVG, comm=evens.get_global_communicator())
        members = list(range(0, n, 2))
        same(t, sum(data(r) for r in members) / len(members), "allreduce_inplace (AVG) on a sub-group")

    bagua.barrier(comm=comm)
    if rank == 0:
        print("all communication primitives match torch.distributed")


if __name__ == "__main__":
    main()
