r"""``bagua_sys_perf``: quick whole-system throughput probe (reference: bagua/script/bagua_sys_perf:1-158) — trains a synthetic
model for a few iterations under the launcher and prints Horovod-style ``Img/sec per GPU`` / ``Total img/sec`` lines that
``bagua_b200.service.autotune_system`` parses."""
from __future__ import annotations

import argparse
import time


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="vgg16")
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--num-iters", type=int, default=5)
    p.add_argument("--num-batches-per-iter", type=int, default=10)
    p.add_argument("--num-warmup-batches", type=int, default=10)
    p.add_argument("--algorithm", default="gradient_allreduce")
    p.add_argument("--cpu", "--no-cuda", dest="cpu", action="store_true", help="gloo + CPU tensors")
    p.add_argument("--fp16-allreduce", action="store_true", default=False,
                   help="communicate 16-bit gradients: the model runs in bf16 (GPU only), so the all-reduce moves half the bytes (the reference's flag, a Horovod leftover, has no effect there)")
    p.add_argument("--use-adasum", action="store_true", default=False, help="accepted for command-line compatibility (Horovod leftover, no effect — as in the reference)")
    args = p.parse_args(argv)

    import torch
    import torch.nn.functional as F

    import bagua_b200 as bagua
    from bagua_b200.models import get_model
    from bagua_b200.parallel.algorithms import Algorithm

    cuda = torch.cuda.is_available() and not args.cpu
    if cuda:
        torch.cuda.set_device(bagua.get_local_rank())
    bagua.init_process_group()
    dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
    model = get_model(args.model).to(dev)
    half = args.fp16_allreduce and cuda
    if half:
        model = model.to(torch.bfloat16)
    opt = torch.optim.SGD(model.parameters(), lr=0.01 * bagua.get_world_size())
    model = model.with_bagua([opt], Algorithm.init(args.algorithm))
    shape = (args.batch_size, 1, 28, 28) if args.model == "mnist" else (args.batch_size, 3, 224, 224)
    classes = 10 if args.model == "mnist" else 1000
    data, target = torch.randn(*shape, device=dev), torch.randint(0, classes, (args.batch_size,), device=dev)
    if half:
        data = data.to(torch.bfloat16)

    def step():
        opt.zero_grad()
        out = model(data)
        loss = F.nll_loss(out.float(), target) if args.model == "mnist" else F.cross_entropy(out.float(), target)
        loss.backward()
        opt.step()

    def sync():
        if cuda:
            torch.cuda.synchronize()

    for _ in range(args.num_warmup_batches):
        step()
    sync()
    speeds = []
    for i in range(args.num_iters):
        t0 = time.time()
        for _ in range(args.num_batches_per_iter):
            step()
        sync()
        s = args.batch_size * args.num_batches_per_iter / (time.time() - t0)
        speeds.append(s)
        if bagua.get_rank() == 0:
            print(f"Iter #{i}: {s:.1f} img/sec per {'GPU' if cuda else 'CPU'}")
    import statistics

    mean = statistics.mean(speeds)
    conf = 1.96 * (statistics.pstdev(speeds) if len(speeds) > 1 else 0.0)
    if bagua.get_rank() == 0:
        n = bagua.get_world_size()
        print(f"Img/sec per GPU: {mean:.1f} +-{conf:.1f}")
        print(f"Total img/sec on {n} GPU(s): {n * mean:.1f} +-{n * conf:.1f}")


if __name__ == "__main__":
    main()
