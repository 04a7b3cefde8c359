"""Synthetic throughput benchmark for every algorithm (reference: examples/benchmark/synthetic_benchmark.py).

    python -m bagua_b200.distributed.launch --nproc_per_node=8 examples/benchmark/synthetic_benchmark.py --model vgg16 --algorithm gradient_allreduce

Prints Horovod-style ``Img/sec per GPU`` / ``Total img/sec`` lines; timing is device-side (CUDA events, max over ranks)."""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bagua_b200 as bagua  # noqa: E402
from bagua_b200.models import get_model  # noqa: E402
from bagua_b200.parallel.algorithms import Algorithm, q_adam  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", default="resnet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--num-warmup-batches", type=int, default=10)
p.add_argument("--num-batches-per-iter", type=int, default=10)
p.add_argument("--num-iters", type=int, default=10)
p.add_argument("--algorithm", default="gradient_allreduce", help="gradient_allreduce, bytegrad, decentralized, low_precision_decentralized, qadam, async")
p.add_argument("--fuse-optimizer", action="store_true")
p.add_argument("--bf16", action="store_true")
p.add_argument("--async-sync-interval", type=int, default=500)
p.add_argument("--async-warmup-steps", type=int, default=100)
p.add_argument("--deterministic", action="store_true", help="fixed seeds + deterministic kernels; prints the final loss (the reference's CI "
               "compares it against a recorded value, .buildkite/scripts/benchmark_master.sh:84)")
p.add_argument("--image-size", type=int, default=224)
p.add_argument("--cpu", "--no-cuda", dest="cpu", action="store_true", help="gloo + CPU tensors (smoke tests)")
p.add_argument("--amp", action="store_true", help="fp32 parameters, forward under torch.autocast (fp16 + GradScaler on GPUs, bf16 on CPUs); the reference's "
               "flag of the same name crashes on an undefined args.scaler (synthetic_benchmark.py:181-186)")
p.add_argument("--log-interval", type=int, default=0, help="also print the loss every N batches")
args = p.parse_args()

cuda = torch.cuda.is_available() and not args.cpu
if cuda:
    torch.cuda.set_device(bagua.get_local_rank())
bagua.init_process_group()
dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
if args.deterministic:
    torch.manual_seed(42)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
else:
    torch.backends.cudnn.benchmark = True
model = get_model(args.model).to(dev)
if args.bf16:
    model = model.to(torch.bfloat16).to(memory_format=torch.channels_last)
if args.algorithm == "qadam":
    optimizer = q_adam.QAdamOptimizer(model.parameters(), lr=0.01 * bagua.get_world_size(), warmup_steps=100)
    algorithm = q_adam.QAdamAlgorithm(optimizer)
else:
    optimizer = torch.optim.SGD(model.parameters(), lr=0.01 * bagua.get_world_size())
    kw = dict(sync_interval_ms=args.async_sync_interval, warmup_steps=args.async_warmup_steps) if args.algorithm == "async" else {}
    algorithm = Algorithm.init(args.algorithm, **kw)
model = model.with_bagua([optimizer], algorithm)
if args.fuse_optimizer:
    optimizer = bagua.contrib.fuse_optimizer(optimizer)

shape = (args.batch_size, 1, 28, 28) if args.model == "mnist" else (args.batch_size, 3, args.image_size, args.image_size)
num_classes = 10 if args.model == "mnist" else 1000
data = torch.randn(*shape, device=dev)
if args.bf16:
    data = data.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
target = torch.randint(0, num_classes, (args.batch_size,), device=dev)


amp_dtype = torch.float16 if cuda else torch.bfloat16
scaler = torch.amp.GradScaler("cuda") if (args.amp and cuda) else None
batches_done = 0


def step():
    global batches_done
    optimizer.zero_grad()
    if args.amp:
        with torch.autocast(dev.type, dtype=amp_dtype):
            loss = F.cross_entropy(model(data).float(), target)
    else:
        loss = F.cross_entropy(model(data).float(), target)
    if scaler is not None:   # gradients are communicated scaled; unscale + inf check happen in scaler.step after the all-reduce
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
    else:
        loss.backward()
        optimizer.fuse_step() if args.fuse_optimizer else optimizer.step()
    batches_done += 1
    if args.log_interval and batches_done % args.log_interval == 0 and bagua.get_rank() == 0:
        print(f"batch {batches_done}: loss {loss.item():.6f}")
    return loss


def timed(n):
    """Device-timed on GPUs (CUDA events, max over ranks); wall clock on the CPU smoke path."""
    if cuda:
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    else:
        import time

        t0 = time.perf_counter()
    for _ in range(n):
        loss = step()
    if cuda:
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
    else:
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3])
    bagua.allreduce_inplace(ms, op=bagua.ReduceOp.MAX)
    return ms.item(), loss


for _ in range(args.num_warmup_batches):
    step()
speeds = []
for i in range(args.num_iters):
    ms, loss = timed(args.num_batches_per_iter)
    speeds.append(args.batch_size * args.num_batches_per_iter / (ms / 1e3))
    if bagua.get_rank() == 0:
        print(f"Iter #{i}: {speeds[-1]:.1f} img/sec per GPU")
if args.algorithm == "async":
    model.bagua_algorithm.abort(model)
if bagua.get_rank() == 0:
    import statistics

    m, c = statistics.mean(speeds), 1.96 * statistics.pstdev(speeds)
    n = bagua.get_world_size()
    print(f"Img/sec per GPU: {m:.1f} +-{c:.1f}")
    print(f"Total img/sec on {n} GPU(s): {n * m:.1f} +-{n * c:.1f}")
    if args.deterministic:
        print(f"Final loss: {loss.item():.6f}")
