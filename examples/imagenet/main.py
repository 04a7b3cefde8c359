"""ImageNet-style training (reference: examples/imagenet/main.py): ResNet-50 / VGG16, bf16 channels-last, any algorithm,
fused optimizers, step LR schedule, top-1/top-5, checkpoint resume, ``--prof`` NVTX ranges + cudaProfilerStart/Stop.

    python -m bagua_b200.distributed.launch --nproc_per_node=8 examples/imagenet/main.py --arch resnet50 --data /data/imagenet
    python -m bagua_b200.distributed.launch --nproc_per_node=8 examples/imagenet/main.py --arch vgg16 --synthetic --fused-shard

``--data DIR`` expects ``DIR/train`` and ``DIR/val`` in torchvision ImageFolder layout (needs torchvision); ``--synthetic``
(the default when no directory is given) generates class-dependent random images so the script runs offline.  Batches
are prefetched to the GPU through pinned memory on a side stream (``bagua_b200.utils.data.DevicePrefetcher``)."""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bagua_b200 as bagua  # noqa: E402
from bagua_b200.models import get_model  # noqa: E402
from bagua_b200.parallel.algorithms import Algorithm, q_adam  # noqa: E402
from bagua_b200.parallel.algorithms.gradient_allreduce import FusedGradientAllReduceAlgorithm, make_sharded_fused_sgd  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--data", default="")
p.add_argument("--synthetic", action="store_true")
p.add_argument("--arch", default="resnet50", choices=["resnet50", "vgg16"])
p.add_argument("--epochs", type=int, default=90)
p.add_argument("--steps-per-epoch", type=int, default=100, help="synthetic data only")
p.add_argument("--batch-size", type=int, default=32, help="per GPU")
p.add_argument("--lr", type=float, default=0.1, help="for a global batch of 256; scaled linearly")
p.add_argument("--momentum", type=float, default=0.9)
p.add_argument("--weight-decay", type=float, default=1e-4)
p.add_argument("--algorithm", default="gradient_allreduce")
p.add_argument("--fused-shard", action="store_true", help="SGD inside the bucket kernels (reduce-scatter → update → all-gather)")
p.add_argument("--fuse-optimizer", action="store_true", help="generic fused optimizer (bagua.contrib.fuse_optimizer)")
p.add_argument("--resume", default="")
p.add_argument("--save", default="")
p.add_argument("--prof", type=int, default=-1, help="profile this many iterations after 10 warm-up steps, then exit")
p.add_argument("--num-classes", type=int, default=1000)
p.add_argument("--image-size", type=int, default=224)
p.add_argument("--print-freq", type=int, default=20)
p.add_argument("--cpu", action="store_true")
args = p.parse_args()

cuda = torch.cuda.is_available() and not args.cpu
if cuda:
    torch.cuda.set_device(bagua.get_local_rank())
bagua.init_process_group()
rank, world = bagua.get_rank(), bagua.get_world_size()
dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
dtype = torch.bfloat16 if cuda else torch.float32
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)


class SyntheticImages(torch.utils.data.Dataset):
    def __init__(self, n, classes, size, seed):
        self.n, self.classes, self.size, self.seed = n, classes, size, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        y = int(torch.randint(0, self.classes, (1,), generator=g))
        x = torch.randn(3, self.size, self.size, generator=g) + (y % 7 - 3) * 0.1  # weak class signal
        return x, y


def loaders():
    if args.data and not args.synthetic:
        import torchvision.datasets as D
        import torchvision.transforms as T

        norm = T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
        train = D.ImageFolder(os.path.join(args.data, "train"), T.Compose([T.RandomResizedCrop(args.image_size), T.RandomHorizontalFlip(), T.ToTensor(), norm]))
        val = D.ImageFolder(os.path.join(args.data, "val"), T.Compose([T.Resize(256), T.CenterCrop(args.image_size), T.ToTensor(), norm]))
        workers = 8
    else:
        train = SyntheticImages(args.steps_per_epoch * args.batch_size * world, args.num_classes, args.image_size, 1)
        val = SyntheticImages(4 * args.batch_size * world, args.num_classes, args.image_size, 2)
        workers = 0
    ts = torch.utils.data.distributed.DistributedSampler(train, num_replicas=world, rank=rank, shuffle=True)
    vs = torch.utils.data.distributed.DistributedSampler(val, num_replicas=world, rank=rank, shuffle=False)
    kw = dict(batch_size=args.batch_size, num_workers=workers, pin_memory=cuda, drop_last=True)
    return torch.utils.data.DataLoader(train, sampler=ts, **kw), torch.utils.data.DataLoader(val, sampler=vs, **kw), ts


model = get_model(args.arch, num_classes=args.num_classes).to(dev).to(dtype)
if cuda:
    model = model.to(memory_format=torch.channels_last)
lr = args.lr * args.batch_size * world / 256.0
if args.algorithm == "qadam":
    optimizer = q_adam.QAdamOptimizer(model.parameters(), lr=1e-3, warmup_steps=100)
    algorithm = q_adam.QAdamAlgorithm(optimizer)
elif args.fused_shard and cuda and world > 1:
    optimizer = make_sharded_fused_sgd(model.parameters(), lr=lr, momentum=args.momentum, weight_decay=args.weight_decay)
    algorithm = FusedGradientAllReduceAlgorithm(optimizer)
else:
    if cuda and not args.fuse_optimizer:
        optimizer = bagua.ops.FusedSGD(model.parameters(), lr=lr, momentum=args.momentum, weight_decay=args.weight_decay)
    else:
        optimizer = torch.optim.SGD(model.parameters(), lr=lr, momentum=args.momentum, weight_decay=args.weight_decay)
    algorithm = Algorithm.init(args.algorithm)
start_epoch = 0
if args.resume and os.path.isfile(args.resume):
    ck = torch.load(args.resume, map_location=dev)
    model.load_state_dict(ck["model"])
    start_epoch = ck["epoch"] + 1
model = model.with_bagua([optimizer], algorithm)
if args.fuse_optimizer:
    optimizer = bagua.contrib.fuse_optimizer(optimizer)
scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda e: 0.1 ** ((e + start_epoch) // 30))
train_loader, val_loader, train_sampler = loaders()


def to_device(x, y):
    x = x.to(dev, non_blocking=True).to(dtype)
    if cuda:
        x = x.contiguous(memory_format=torch.channels_last)
    return x, y.to(dev, non_blocking=True)


def accuracy(out, y, ks=(1, 5)):
    _, pred = out.topk(max(ks), 1)
    hit = pred.eq(y.view(-1, 1))
    return [hit[:, :k].any(1).float().mean() for k in ks]


def nvtx(name):
    return torch.cuda.nvtx.range(name) if cuda and args.prof >= 0 else __import__("contextlib").nullcontext()


def train(epoch):
    model.train()
    train_sampler.set_epoch(epoch)
    if args.algorithm == "async":
        model.bagua_algorithm.resume(model)
    t0, seen = time.time(), 0
    for i, (x, y) in enumerate(train_loader):
        if args.prof >= 0 and i == 10 and cuda:
            torch.cuda.cudart().cudaProfilerStart()
        with nvtx(f"iter{i}"):
            x, y = to_device(x, y)
            optimizer.zero_grad()
            with nvtx("forward"):
                out = model(x)
                loss = F.cross_entropy(out.float(), y)
            with nvtx("backward"):
                loss.backward()
            with nvtx("step"):
                optimizer.fuse_step() if args.fuse_optimizer else optimizer.step()
        seen += x.shape[0]
        if i % args.print_freq == 0 and rank == 0:
            print(f"epoch {epoch} [{i}/{len(train_loader)}] loss {loss.item():.4f} lr {optimizer.param_groups[0]['lr']:.4g} "
                  f"{seen * world / (time.time() - t0):.0f} img/s", flush=True)
        if args.prof >= 0 and i == 10 + args.prof:
            if cuda:
                torch.cuda.cudart().cudaProfilerStop()
            return False
    if args.algorithm == "async":
        model.bagua_algorithm.abort(model)
    return True


@torch.no_grad()
def validate():
    model.eval()
    tot = torch.zeros(3, device=dev)
    for x, y in val_loader:
        x, y = to_device(x, y)
        out = model(x).float()
        a1, a5 = accuracy(out, y)
        tot += torch.stack([a1, a5, torch.ones((), device=dev)])
    bagua.allreduce_inplace(tot, op=bagua.ReduceOp.SUM)
    return (tot[0] / tot[2]).item(), (tot[1] / tot[2]).item()


for epoch in range(start_epoch, args.epochs):
    if not train(epoch):
        break
    scheduler.step()
    top1, top5 = validate()
    if rank == 0:
        print(f"epoch {epoch}: top1 {top1 * 100:.2f} top5 {top5 * 100:.2f}", flush=True)
        if args.save:
            torch.save({"model": model.state_dict(), "epoch": epoch}, args.save + ".tmp")
            os.replace(args.save + ".tmp", args.save)
