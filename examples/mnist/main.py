"""MNIST-shaped training with any algorithm (reference: examples/mnist/main.py). Uses synthetic digits when torchvision's
dataset is not on disk, so it runs offline and on CPU (gloo) as well as on GPUs.

    python -m bagua_b200.distributed.launch --nproc_per_node=2 examples/mnist/main.py --algorithm bytegrad --epochs 1"""
import argparse

import torch
import torch.nn.functional as F

import bagua_b200 as bagua
from bagua_b200.models import MnistNet
from bagua_b200.parallel.algorithms import Algorithm, q_adam

p = argparse.ArgumentParser()
p.add_argument("--batch-size", type=int, default=64)
p.add_argument("--epochs", type=int, default=1)
p.add_argument("--steps-per-epoch", type=int, default=50)
p.add_argument("--lr", type=float, default=0.05)
p.add_argument("--algorithm", default="gradient_allreduce")
p.add_argument("--fuse-optimizer", action="store_true")
p.add_argument("--cpu", action="store_true")
args = p.parse_args()

cuda = torch.cuda.is_available() and not args.cpu
if cuda:
    torch.cuda.set_device(bagua.get_local_rank())
bagua.init_process_group()
dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
torch.manual_seed(1)
model = MnistNet().to(dev)
if args.algorithm == "qadam":
    optimizer = q_adam.QAdamOptimizer(model.parameters(), lr=1e-3, warmup_steps=20)
    algorithm = q_adam.QAdamAlgorithm(optimizer)
else:
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr)
    algorithm = Algorithm.init(args.algorithm, **({"sync_interval_ms": 50} if args.algorithm == "async" else {}))
model = model.with_bagua([optimizer], algorithm)
if args.fuse_optimizer:
    optimizer = bagua.contrib.fuse_optimizer(optimizer)
g = torch.Generator().manual_seed(bagua.get_rank())
for epoch in range(args.epochs):
    if args.algorithm == "async":
        model.bagua_algorithm.resume(model)
    for it in range(args.steps_per_epoch):
        y = torch.randint(0, 10, (args.batch_size,), generator=g)
        x = torch.randn(args.batch_size, 1, 28, 28, generator=g) + y.view(-1, 1, 1, 1).float() * 0.3  # class-dependent signal
        x, y = x.to(dev), y.to(dev)
        optimizer.zero_grad()
        loss = F.nll_loss(model(x), y)
        loss.backward()
        optimizer.fuse_step() if args.fuse_optimizer else optimizer.step()
    if args.algorithm == "async":
        model.bagua_algorithm.abort(model)
    if bagua.get_rank() == 0:
        print(f"epoch {epoch}: loss {loss.item():.4f}")
