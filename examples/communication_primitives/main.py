"""Every blocking collective next to torch.distributed (reference: examples/communication_primitives/main.py)."""
import torch
import torch.distributed as dist

import bagua_b200 as bagua

cuda = torch.cuda.is_available()
if cuda:
    torch.cuda.set_device(bagua.get_local_rank())
bagua.init_process_group()
dev = "cuda" if cuda else "cpu"
rank, n = bagua.get_rank(), bagua.get_world_size()

x = torch.rand(1000, device=dev)
a, b = x.clone(), x.clone()
bagua.allreduce_inplace(a)
dist.all_reduce(b)
assert torch.allclose(a, b), "allreduce"
a, b = x.clone(), x.clone()
bagua.broadcast(a, 0)
dist.broadcast(b, 0)
assert torch.equal(a, b), "broadcast"
out_a, out_b = torch.zeros(1000 * n, device=dev), [torch.zeros(1000, device=dev) for _ in range(n)]
bagua.allgather(x, out_a)
dist.all_gather(out_b, x)
assert torch.equal(out_a, torch.cat(out_b)), "allgather"
send = torch.arange(n, dtype=torch.float32, device=dev) + rank * n
ra, rb = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
bagua.alltoall(send, ra)
if cuda:
    dist.all_to_all_single(rb, send)
    assert torch.equal(ra, rb), "alltoall"
big = torch.rand(n * 10, device=dev)
o = torch.zeros(10, device=dev)
bagua.reduce_scatter(big, o)
ref = big.clone()
dist.all_reduce(ref)
assert torch.allclose(o, ref[rank * 10:(rank + 1) * 10]), "reduce_scatter"
bagua.barrier()
if rank == 0:
    print("all communication primitives match torch.distributed")
