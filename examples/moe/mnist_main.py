"""Mixture-of-experts CNN on MNIST-shaped data, any algorithm, optional checkpoint round trip
(same command line as the reference's examples/moe/mnist_main.py: --batch-size --epochs --lr --gamma --log-interval --save-model
--algorithm --async-sync-interval --set-deterministic --num-local-experts).

    python -m bagua_b200.distributed.launch --nproc_per_node=2 examples/moe/mnist_main.py --num-local-experts 2 --epochs 2

There is no network in the build image, so the "dataset" is a deterministic synthetic one: class-dependent blobs with the MNIST
tensor shapes (``--steps-per-epoch`` batches per epoch).  With ``--save-model`` the model and optimizer are saved through
``bagua_b200.checkpoint`` after every epoch and loaded back before the next one — the final loss must not change (the reference's CI
compares it exactly, .buildkite/scripts/benchmark_master.sh:137-151; ``scripts/ci/benchmark_ci.sh`` does the same here)."""
import argparse
import logging
import os
import sys
import tempfile

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bagua_b200 as bagua  # noqa: E402
from bagua_b200.checkpoint import load_checkpoint, save_checkpoint  # noqa: E402
from bagua_b200.parallel.algorithms import Algorithm, q_adam  # noqa: E402


def parse():
    p = argparse.ArgumentParser(description="bagua_b200 MoE MNIST example")
    p.add_argument("--batch-size", type=int, default=64)
    p.add_argument("--test-batch-size", type=int, default=1000)
    p.add_argument("--epochs", type=int, default=14)
    p.add_argument("--lr", type=float, default=1.0)
    p.add_argument("--gamma", type=float, default=0.7, help="learning-rate decay per epoch")
    p.add_argument("--log-interval", type=int, default=10)
    p.add_argument("--save-model", action="store_true", default=False, help="checkpoint after every epoch and reload before the next")
    p.add_argument("--algorithm", default="gradient_allreduce", help="gradient_allreduce, bytegrad, decentralized, low_precision_decentralized, qadam, async")
    p.add_argument("--async-sync-interval", type=int, default=500)
    p.add_argument("--set-deterministic", action="store_true", default=False)
    p.add_argument("--num-local-experts", type=int, default=0, help="experts per GPU (0: dense model)")
    p.add_argument("--steps-per-epoch", type=int, default=15, help="synthetic batches per epoch")
    p.add_argument("--save-dir", default=None)
    p.add_argument("--cpu", action="store_true")
    return p.parse_args()


class Net(nn.Module):
    def __init__(self, num_local_experts: int):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(1, 32, 3, 1), nn.Conv2d(32, 64, 3, 1)
        self.drop1, self.drop2 = nn.Dropout(0.25), nn.Dropout(0.5)
        self.fc1 = nn.Linear(9216, 128)
        self.moe = bagua.moe.MoE(128, nn.Linear(128, 128), num_local_experts, k=2) if num_local_experts > 0 else None
        self.fc2 = nn.Linear(128, 10)

    def forward(self, x):
        x = self.drop1(F.max_pool2d(F.relu(self.conv2(F.relu(self.conv1(x)))), 2))
        x = self.drop2(F.relu(self.fc1(torch.flatten(x, 1))))
        aux = x.new_zeros(())
        if self.moe is not None:
            x, aux, _ = self.moe(x)
        return F.log_softmax(self.fc2(x), dim=1), aux


def batches(epoch: int, args, rank: int):
    g = torch.Generator().manual_seed(1000 * epoch + rank)
    for _ in range(args.steps_per_epoch):
        y = torch.randint(0, 10, (args.batch_size,), generator=g)
        yield torch.randn(args.batch_size, 1, 28, 28, generator=g) + y.view(-1, 1, 1, 1).float() * 0.3, y


def main():
    args = parse()
    logging.basicConfig(level=logging.INFO, format="%(message)s", stream=sys.stdout)
    cuda = torch.cuda.is_available() and not args.cpu
    if cuda:
        torch.cuda.set_device(bagua.get_local_rank())
    bagua.init_process_group()
    rank = bagua.get_rank()
    dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
    if args.set_deterministic:
        print("set_deterministic: True")
        torch.manual_seed(0)
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
        torch.set_printoptions(precision=10)
    else:
        torch.manual_seed(rank)
    model = Net(args.num_local_experts).to(dev)
    if args.algorithm == "qadam":
        optimizer = q_adam.QAdamOptimizer(model.parameters(), lr=args.lr * 1e-3, warmup_steps=10)
        algorithm = q_adam.QAdamAlgorithm(optimizer)
    else:
        optimizer = torch.optim.Adadelta(model.parameters(), lr=args.lr)
        kw = {"sync_interval_ms": args.async_sync_interval} if args.algorithm == "async" else {}
        algorithm = Algorithm.init(args.algorithm, **kw)
    model = model.with_bagua([optimizer], algorithm)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=args.gamma)
    save_dir = None
    if args.save_model:
        save_dir = bagua.broadcast_object((args.save_dir or tempfile.mkdtemp()) if rank == 0 else None, 0)
    loss = None
    for epoch in range(1, args.epochs + 1):
        if args.algorithm == "async":
            model.bagua_algorithm.resume(model)
        model.train()
        for i, (x, y) in enumerate(batches(epoch, args, rank)):
            optimizer.zero_grad()
            out, aux = model(x.to(dev))
            loss = F.nll_loss(out, y.to(dev)) + 0.01 * aux
            loss.backward()
            optimizer.step()
            if i % args.log_interval == 0 and rank == 0:
                logging.info("Train Epoch: {} [{}/{}]\tLoss: {:.6f}".format(epoch, i * args.batch_size, args.steps_per_epoch * args.batch_size, loss.item()))
        if args.algorithm == "async":
            model.bagua_algorithm.abort(model)
        scheduler.step()
        if save_dir is not None:   # experts are saved per expert-parallel rank, the dense part once (bagua_b200.checkpoint)
            save_checkpoint(epoch, save_dir, model, optimizer, scheduler)
            assert load_checkpoint(save_dir, model, optimizer, scheduler) == epoch
    if rank == 0:
        logging.info("Final Loss: {:.6f}".format(loss.item()))


if __name__ == "__main__":
    main()
