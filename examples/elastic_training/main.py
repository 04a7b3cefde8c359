"""Elastic MNIST training: workers may join or leave between rendezvous rounds, every (re)started gang resumes from the last
epoch checkpoint — the command line of the reference's examples/elastic_training/main.py:92-160 (--batch-size --test-batch-size
--epochs --lr --gamma --seed --log-interval --save-model --algorithm --checkpoint_path --data_dir).

Node 1 (192.168.1.1 with free port 1234); nodes 2-4 join later with the very same command:

    python -m bagua_b200.distributed.run --nnodes=1:4 --nproc_per_node=8 --rdzv_id=JOB_ID --rdzv_backend=c10d \
        --rdzv_endpoint=192.168.1.1:1234 examples/elastic_training/main.py --checkpoint_path /shared/mnist_elastic.pt

Restart semantics are torch elastic's: when membership changes all workers are restarted, ``bagua_b200.init_process_group``
rendezvouses the new gang (a restarted attempt gets its own key prefix in the agent's store, communication.py), and this script
reloads model / optimizer / scheduler and continues with the epoch that was interrupted.  ``--batch-size`` is the global batch,
so the per-rank batch and the DistributedSampler follow the new world size.  Compared with the reference's example the
checkpoint is written atomically (a worker killed mid-write must not leave a torn file for the next gang) and carries the
scheduler.  Training loop, data and model are the MNIST example's (examples/mnist/main.py)."""
import argparse
import importlib.util
import logging
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import bagua_b200 as bagua  # noqa: E402
from bagua_b200.models import MnistNet  # noqa: E402

_spec = importlib.util.spec_from_file_location("bagua_example_mnist", os.path.join(os.path.dirname(HERE), "mnist", "main.py"))
mnist = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mnist)


def parse():
    p = argparse.ArgumentParser(description="bagua_b200 elastic training example")
    p.add_argument("--batch-size", type=int, default=64, metavar="N", help="global training batch")
    p.add_argument("--test-batch-size", type=int, default=1000, metavar="N")
    p.add_argument("--epochs", type=int, default=14, metavar="N")
    p.add_argument("--lr", type=float, default=1.0, metavar="LR")
    p.add_argument("--gamma", type=float, default=0.7, metavar="M")
    p.add_argument("--seed", type=int, default=1, metavar="S")
    p.add_argument("--log-interval", type=int, default=10, metavar="N")
    p.add_argument("--save-model", action="store_true", default=False)
    p.add_argument("--algorithm", default="gradient_allreduce", help="gradient_allreduce, bytegrad, decentralized, low_precision_decentralized, qadam")
    p.add_argument("--checkpoint_path", default="/tmp/bagua_b200_mnist_elastic.pt")
    p.add_argument("--data_dir", "--data-dir", dest="data_dir", default="../data")
    p.add_argument("--steps-per-epoch", type=int, default=50, help="synthetic data only: global batches per epoch")
    p.add_argument("--cpu", action="store_true")
    args = p.parse_args()
    args.fuse_optimizer, args.async_sync_interval = False, 500   # knobs of the shared MNIST helpers this example does not expose
    return args


def save_atomically(state, path):
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(state, tmp)
    os.replace(tmp, path)


def main():
    args = parse()
    torch.manual_seed(args.seed)
    cuda = torch.cuda.is_available() and not args.cpu
    if cuda:
        torch.cuda.set_device(bagua.get_local_rank())
    bagua.init_process_group()
    rank, world = bagua.get_rank(), bagua.get_world_size()
    logging.basicConfig(format="%(levelname)s:%(message)s", level=logging.INFO if rank == 0 else logging.ERROR, stream=sys.stdout)
    dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
    logging.info("gang of %d workers, restart count %s", world, os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"))

    train_set, test_set, kind = mnist.datasets_for(args)
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, num_replicas=world, rank=rank)
    loader_kw = {"num_workers": 1, "pin_memory": True} if cuda else {}
    train_loader = torch.utils.data.DataLoader(train_set, batch_size=max(args.batch_size // world, 1), sampler=sampler, **loader_kw)
    test_loader = torch.utils.data.DataLoader(test_set, batch_size=args.test_batch_size, **loader_kw)

    model = MnistNet().to(dev)
    optimizer, algorithm = mnist.make_algorithm(args, model)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=args.gamma)
    start_epoch = 1
    if os.path.exists(args.checkpoint_path):   # every worker of a (re)started gang resumes from the same snapshot
        ckpt = torch.load(args.checkpoint_path, map_location=dev)
        model.load_state_dict(ckpt["model_state_dict"])
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if "scheduler_state_dict" in ckpt:
            scheduler.load_state_dict(ckpt["scheduler_state_dict"])
        start_epoch = ckpt["epoch"]
        logging.info("resumed from %s: next epoch %d", args.checkpoint_path, start_epoch)
    # with_bagua broadcasts rank 0's parameters and optimizer state, so a worker that raced the checkpoint write still starts equal
    model = model.with_bagua([optimizer], algorithm)

    loss = None
    for epoch in range(start_epoch, args.epochs + 1):
        sampler.set_epoch(epoch)
        loss = mnist.train_one_epoch(args, model, train_loader, optimizer, epoch, dev)
        mnist.evaluate(model, test_loader, dev)
        scheduler.step()
        if rank == 0:
            save_atomically({"epoch": epoch + 1, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                             "scheduler_state_dict": scheduler.state_dict()}, args.checkpoint_path)
    if rank == 0:
        print(f"done after epoch {args.epochs} (restart count {os.environ.get('TORCHELASTIC_RESTART_COUNT', '0')}"
              + (f", final loss {loss.item():.6f})" if loss is not None else ", nothing left to train)"))
        if args.save_model:
            torch.save(model.state_dict(), "mnist_cnn_elastic.pt")


if __name__ == "__main__":
    main()
