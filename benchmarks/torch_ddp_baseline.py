"""Same-box baseline for bench.py: the identical VGG16 / ResNet-50 step on stock PyTorch — ``torch.nn.parallel.DistributedDataParallel``
(NCCL bucketed all-reduce), ``torch.optim.SGD``, eager bias/ReLU/pool modules — timed exactly like bench.py (CUDA events over K
steps after W warm-up steps, max over ranks, one JSON line from rank 0).

The reference itself cannot be installed offline; this is what its schedule reduces to without its Rust core:
bucketed NCCL all-reduce overlapped with backward.  Not used by the driver; numbers go to profiles/.

    python benchmarks/torch_ddp_baseline.py --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 benchmarks/torch_ddp_baseline.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bagua_b200.models import get_model  # noqa: E402  (architecture only; the fused epilogues are switched off below)

p = argparse.ArgumentParser()
p.add_argument("--steps", type=int, default=30)
p.add_argument("--warmup", type=int, default=5)
p.add_argument("--model", default="vgg16", choices=["vgg16", "resnet50"])
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--momentum", type=float, default=0.0)
p.add_argument("--bucket-cap-mb", type=int, default=25)
args = p.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl")
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234 + rank)
model = get_model(args.model)
if hasattr(model, "fuse_epilogues"):
    model.fuse_epilogues = False  # stock eager Conv2d → ReLU → MaxPool2d modules, no bagua_b200 kernels on this arm
model = model.to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
if world > 1:
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=args.bucket_cap_mb, gradient_as_bucket_view=True)
opt = torch.optim.SGD(model.parameters(), lr=0.01 * world, momentum=args.momentum)
x = torch.randn(args.batch_size, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (args.batch_size,), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(x).float(), y)
    loss.backward()
    opt.step()
    return loss


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(max(args.warmup, 3)):
    step()
sync()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(args.steps):
    loss = step()
e.record()
torch.cuda.synchronize()
ms = torch.tensor([s.elapsed_time(e)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"impl": "torch_ddp", "metric": f"{args.model} synthetic-ImageNet training throughput (torch DDP + torch.optim.SGD)",
                      "value": args.batch_size * world * args.steps / (ms.item() / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms.item() / args.steps, "dtype": "bf16", "data": "synthetic",
                      "loss_finite": bool(torch.isfinite(loss.detach().float()).item())}))
if world > 1:
    dist.destroy_process_group()
