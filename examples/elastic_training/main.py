"""Elastic training (reference: examples/elastic_training/main.py): restart-all semantics of torch elastic + manual resume.

    python -m bagua_b200.distributed.run --nnodes=1:4 --nproc_per_node=8 --rdzv_id=job1 --rdzv_backend=c10d --rdzv_endpoint=host:29400 \
        examples/elastic_training/main.py --ckpt /shared/ckpt.pt"""
import argparse
import os

import torch
import torch.nn.functional as F

import bagua_b200 as bagua
from bagua_b200.models import MnistNet
from bagua_b200.parallel.algorithms import gradient_allreduce

p = argparse.ArgumentParser()
p.add_argument("--ckpt", default="/tmp/bagua_elastic_ckpt.pt")
p.add_argument("--steps", type=int, default=200)
p.add_argument("--cpu", action="store_true")
args = p.parse_args()
cuda = torch.cuda.is_available() and not args.cpu
if cuda:
    torch.cuda.set_device(bagua.get_local_rank())
bagua.init_process_group()
dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
model = MnistNet().to(dev)
optimizer = torch.optim.SGD(model.parameters(), lr=0.05)
start = 0
if os.path.isfile(args.ckpt):  # every worker of a restarted job resumes from the last snapshot
    state = torch.load(args.ckpt, map_location=dev)
    model.load_state_dict(state["model"])
    optimizer.load_state_dict(state["optimizer"])
    start = state["step"] + 1
model = model.with_bagua([optimizer], gradient_allreduce.GradientAllReduceAlgorithm())
for step in range(start, args.steps):
    x, y = torch.randn(32, 1, 28, 28, device=dev), torch.randint(0, 10, (32,), device=dev)
    optimizer.zero_grad()
    F.nll_loss(model(x), y).backward()
    optimizer.step()
    if step % 50 == 0 and bagua.get_rank() == 0:
        tmp = args.ckpt + ".tmp"
        torch.save({"model": model.state_dict(), "optimizer": optimizer.state_dict(), "step": step}, tmp)
        os.replace(tmp, args.ckpt)
if bagua.get_rank() == 0:
    print(f"done at step {args.steps} (restart count {os.environ.get('TORCHELASTIC_RESTART_COUNT', '0')})")
