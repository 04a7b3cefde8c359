"""MoE on MNIST-shaped data with checkpoint save+load (reference: examples/moe/mnist_main.py).

    python -m bagua_b200.distributed.launch --nproc_per_node=2 examples/moe/mnist_main.py --num-local-experts 2"""
import argparse
import tempfile

import torch
import torch.nn as nn
import torch.nn.functional as F

import bagua_b200 as bagua
from bagua_b200.checkpoint import load_checkpoint, save_checkpoint
from bagua_b200.parallel.algorithms import gradient_allreduce

p = argparse.ArgumentParser()
p.add_argument("--num-local-experts", type=int, default=2)
p.add_argument("--steps", type=int, default=30)
p.add_argument("--save-dir", default=None)
p.add_argument("--cpu", action="store_true")
args = p.parse_args()
cuda = torch.cuda.is_available() and not args.cpu
if cuda:
    torch.cuda.set_device(bagua.get_local_rank())
bagua.init_process_group()
dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(1, 32, 3, 1), nn.Conv2d(32, 64, 3, 1)
        self.fc1 = nn.Linear(9216, 128)
        self.moe = bagua.moe.MoE(128, nn.Linear(128, 128), args.num_local_experts, k=2)
        self.fc2 = nn.Linear(128, 10)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.conv2(F.relu(self.conv1(x)))), 2)
        x = F.relu(self.fc1(torch.flatten(x, 1)))
        x, l_aux, _ = self.moe(x)
        return F.log_softmax(self.fc2(x), dim=1), l_aux


torch.manual_seed(bagua.get_rank())
model = Net().to(dev)
optimizer = torch.optim.Adadelta(model.parameters(), lr=1.0)
model = model.with_bagua([optimizer], gradient_allreduce.GradientAllReduceAlgorithm())
g = torch.Generator().manual_seed(bagua.get_rank())
for it in range(args.steps):
    y = torch.randint(0, 10, (64,), generator=g)
    x = (torch.randn(64, 1, 28, 28, generator=g) + y.view(-1, 1, 1, 1).float() * 0.3).to(dev)
    optimizer.zero_grad()
    out, l_aux = model(x)
    loss = F.nll_loss(out, y.to(dev)) + 0.01 * l_aux
    loss.backward()
    optimizer.step()
d = args.save_dir or tempfile.mkdtemp() if bagua.get_rank() == 0 else args.save_dir
d = bagua.broadcast_object(d, 0)
save_checkpoint(args.steps, d, model, optimizer)
assert load_checkpoint(d, model, optimizer) == args.steps
if bagua.get_rank() == 0:
    print(f"final loss {loss.item():.6f}; checkpoint round-trip ok in {d}")
