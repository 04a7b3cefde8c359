"""MNIST with any of the six algorithms — same command line as the reference's examples/mnist/main.py:97-168
(--batch-size --test-batch-size --epochs --lr --gamma --log-interval --save-model --algorithm --async-sync-interval
--set-deterministic --fuse-optimizer) and the same training recipe (conv32-conv64-fc128-fc10, Adadelta, StepLR(1, gamma), a
DistributedSampler over the training set, ``--batch-size`` is the GLOBAL batch, test-set accuracy after every epoch, the async
algorithm resumed / aborted around each epoch, reference :258-277).

    python -m bagua_b200.distributed.launch --nproc_per_node=8 examples/mnist/main.py --algorithm bytegrad --epochs 2

The real dataset is used when torchvision finds it under ``--data-dir`` (no download: the build image has no network);
otherwise a deterministic synthetic stand-in with the MNIST shapes — class-dependent blobs that the network learns in a few
steps — of ``--steps-per-epoch`` global batches.  ``--cpu`` runs the same script on the gloo backend."""
import argparse
import logging
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bagua_b200 as bagua  # noqa: E402
from bagua_b200.models import MnistNet  # noqa: E402
from bagua_b200.parallel.algorithms import Algorithm, q_adam  # noqa: E402


def parse():
    p = argparse.ArgumentParser(description="bagua_b200 MNIST example")
    p.add_argument("--batch-size", type=int, default=64, metavar="N", help="global training batch (split over the ranks)")
    p.add_argument("--test-batch-size", type=int, default=1000, metavar="N")
    p.add_argument("--epochs", type=int, default=14, metavar="N")
    p.add_argument("--lr", type=float, default=1.0, metavar="LR")
    p.add_argument("--gamma", type=float, default=0.7, metavar="M", help="learning-rate decay per epoch")
    p.add_argument("--log-interval", type=int, default=10, metavar="N")
    p.add_argument("--save-model", action="store_true", default=False, help="rank 0 writes mnist_cnn.pt at the end")
    p.add_argument("--algorithm", default="gradient_allreduce", help="gradient_allreduce, bytegrad, decentralized, low_precision_decentralized, qadam, async")
    p.add_argument("--async-sync-interval", type=int, default=500, help="model averaging interval (ms) of the async algorithm")
    p.add_argument("--set-deterministic", action="store_true", default=False)
    p.add_argument("--fuse-optimizer", action="store_true", default=False)
    p.add_argument("--data-dir", default="../data")
    p.add_argument("--steps-per-epoch", type=int, default=50, help="synthetic data only: global batches per epoch")
    p.add_argument("--cpu", action="store_true")
    return p.parse_args()


class SyntheticDigits(torch.utils.data.Dataset):
    """``n`` images of shape [1, 28, 28]: unit noise plus a class-dependent offset, generated per index (no storage)."""

    def __init__(self, n: int, seed: int):
        self.n, self.seed = n, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        y = int(torch.randint(0, 10, (1,), generator=g))
        return torch.randn(1, 28, 28, generator=g) + 0.3 * y, y


def datasets_for(args):
    try:
        from torchvision import datasets, transforms

        tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.1307,), (0.3081,))])
        return datasets.MNIST(args.data_dir, train=True, download=False, transform=tf), datasets.MNIST(args.data_dir, train=False, download=False, transform=tf), "MNIST"
    except Exception:  # noqa: BLE001 - torchvision missing or the files are not on disk
        return SyntheticDigits(args.steps_per_epoch * args.batch_size, 1), SyntheticDigits(max(args.test_batch_size, 200), 2), "synthetic"


def make_algorithm(args, model):
    """The optimizer/algorithm pair of the reference's example (:222-257): Adadelta for every family except QAdam, which brings
    its own optimizer."""
    if args.algorithm == "qadam":
        optimizer = q_adam.QAdamOptimizer(model.parameters(), lr=args.lr * 1e-3, warmup_steps=100)
        return optimizer, q_adam.QAdamAlgorithm(optimizer)
    optimizer = torch.optim.Adadelta(model.parameters(), lr=args.lr)
    kw = {"sync_interval_ms": args.async_sync_interval} if args.algorithm == "async" else {}
    return optimizer, Algorithm.init(args.algorithm, **kw)


def train_one_epoch(args, model, loader, optimizer, epoch, dev):
    model.train()
    loss = None
    for i, (x, y) in enumerate(loader):
        x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
        optimizer.zero_grad()
        loss = F.nll_loss(model(x), y)
        loss.backward()
        optimizer.fuse_step() if args.fuse_optimizer else optimizer.step()
        if i % args.log_interval == 0:
            logging.info("Train Epoch: {} [{}/{} ({:.0f}%)]\tLoss: {:.6f}".format(epoch, i * len(x), len(loader.dataset) // bagua.get_world_size(),
                                                                                  100.0 * i / max(len(loader), 1), loss.item()))
    return loss


@torch.no_grad()
def evaluate(model, loader, dev):
    model.eval()
    total, correct, n = 0.0, 0, 0
    for x, y in loader:
        x, y = x.to(dev), y.to(dev)
        out = model(x)
        total += F.nll_loss(out, y, reduction="sum").item()
        correct += int((out.argmax(1) == y).sum())
        n += len(y)
    logging.info("\nTest set: Average loss: {:.4f}, Accuracy: {}/{} ({:.0f}%)\n".format(total / max(n, 1), correct, n, 100.0 * correct / max(n, 1)))
    return correct / max(n, 1)


def main():
    args = parse()
    cuda = torch.cuda.is_available() and not args.cpu
    if args.set_deterministic:
        print("set_deterministic: True")
        np.random.seed(666)
        random.seed(666)
        torch.backends.cudnn.benchmark, torch.backends.cudnn.deterministic = False, True
        torch.manual_seed(666)
        if cuda:
            torch.cuda.manual_seed_all(666 + bagua.get_rank())
        torch.set_printoptions(precision=10)
    if cuda:
        torch.cuda.set_device(bagua.get_local_rank())
    bagua.init_process_group()
    rank, world = bagua.get_rank(), bagua.get_world_size()
    logging.basicConfig(format="%(levelname)s:%(message)s", level=logging.INFO if rank == 0 else logging.ERROR, stream=sys.stdout)
    dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")

    train_set, test_set, kind = datasets_for(args)
    logging.info("data: %s (%d training samples)", kind, len(train_set))
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, num_replicas=world, rank=rank)
    loader_kw = {"num_workers": 1, "pin_memory": True} if cuda else {}
    train_loader = torch.utils.data.DataLoader(train_set, batch_size=max(args.batch_size // world, 1), sampler=sampler, **loader_kw)
    test_loader = torch.utils.data.DataLoader(test_set, batch_size=args.test_batch_size, **loader_kw)

    model = MnistNet().to(dev)
    optimizer, algorithm = make_algorithm(args, model)
    # the generic fused optimizer flattens parameters itself, so the engine must not (reference :259-263)
    model = model.with_bagua([optimizer], algorithm, do_flatten=not args.fuse_optimizer)
    if args.fuse_optimizer:
        optimizer = bagua.contrib.fuse_optimizer(optimizer)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=args.gamma)

    loss = None
    for epoch in range(1, args.epochs + 1):
        sampler.set_epoch(epoch)
        if args.algorithm == "async":
            model.bagua_algorithm.resume(model)
        loss = train_one_epoch(args, model, train_loader, optimizer, epoch, dev)
        if args.algorithm == "async":
            model.bagua_algorithm.abort(model)
        evaluate(model, test_loader, dev)
        scheduler.step()
    if rank == 0 and loss is not None:
        print(f"final loss {loss.item():.6f}")
    if args.save_model and rank == 0:
        torch.save(model.state_dict(), "mnist_cnn.pt")


if __name__ == "__main__":
    main()
