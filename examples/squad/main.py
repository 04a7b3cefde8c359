"""SQuAD-style question-answering fine-tune of BERT-large (reference: examples/squad/main.py) with any algorithm.

    python -m bagua_b200.distributed.launch --nproc_per_node=8 examples/squad/main.py --algorithm bytegrad
    python -m bagua_b200.distributed.launch --nproc_per_node=8 examples/squad/main.py --features train_features.pt --init bert_large.pt

``--features FILE`` is a ``torch.save``d dict of pre-tokenised tensors (``input_ids, token_type_ids, attention_mask,
start_positions, end_positions`` — what HuggingFace's ``squad_convert_examples_to_features`` produces); without it the
script synthesises SQuAD-shaped features (sequence length 384, a question segment followed by a context segment, an answer
span inside the context) so it runs offline.  ``--init FILE`` loads a state dict (e.g. converted pretrained weights).
Optimiser: AdamW with linear warm-up/decay — ``FusedAdam`` (one kernel per bucket arena), the in-kernel sharded Adam
(``--fused-shard``) or the generic ``fuse_optimizer`` wrapper.  Evaluation reports exact-match of the predicted span."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bagua_b200 as bagua  # noqa: E402
from bagua_b200 import models  # noqa: E402
from bagua_b200.contrib import LoadBalancingDistributedSampler  # noqa: E402
from bagua_b200.parallel.algorithms import Algorithm, q_adam  # noqa: E402
from bagua_b200.parallel.algorithms.gradient_allreduce import FusedGradientAllReduceAlgorithm, make_sharded_fused_adam  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--features", default="")
p.add_argument("--init", default="")
p.add_argument("--algorithm", default="gradient_allreduce")
p.add_argument("--epochs", type=int, default=2)
p.add_argument("--batch-size", type=int, default=8, help="per GPU")
p.add_argument("--max-seq-length", type=int, default=384)
p.add_argument("--learning-rate", type=float, default=3e-5)
p.add_argument("--warmup-ratio", type=float, default=0.1)
p.add_argument("--weight-decay", type=float, default=0.01)
p.add_argument("--num-synthetic", type=int, default=2048)
p.add_argument("--fused-shard", action="store_true")
p.add_argument("--fuse-optimizer", action="store_true")
p.add_argument("--load-balance", action="store_true", help="LoadBalancingDistributedSampler keyed on the number of non-pad tokens")
p.add_argument("--save-dir", default="")
p.add_argument("--tiny", action="store_true", help="2-layer model for smoke tests")
p.add_argument("--cpu", action="store_true")
p.add_argument("--print-freq", type=int, default=20)
args = p.parse_args()

cuda = torch.cuda.is_available() and not args.cpu
if cuda:
    torch.cuda.set_device(bagua.get_local_rank())
bagua.init_process_group()
rank, world = bagua.get_rank(), bagua.get_world_size()
dev = torch.device("cuda", bagua.get_local_rank()) if cuda else torch.device("cpu")
dtype = torch.bfloat16 if cuda else torch.float32
torch.manual_seed(42)

cfg = models.bert_large_config() if not args.tiny else models.BertConfig(vocab_size=1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                                                         intermediate_size=128, max_position_embeddings=args.max_seq_length)


def synthetic_features(n, seq, vocab):
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(1000 if vocab > 2000 else 10, vocab, (n, seq), generator=g)
    qlen = torch.randint(8, 32, (n,), generator=g)
    total = torch.randint(seq // 2, seq + 1, (n,), generator=g)
    ar = torch.arange(seq).unsqueeze(0)
    mask = (ar < total.unsqueeze(1)).long()
    types = ((ar >= qlen.unsqueeze(1)) & (ar < total.unsqueeze(1))).long()
    ids = ids * mask
    start = (qlen + (torch.rand(n, generator=g) * (total - qlen - 4).clamp(min=1)).long()).clamp(max=seq - 2)
    end = torch.minimum(start + torch.randint(0, 4, (n,), generator=g), total - 1)
    # make the task learnable: mark the answer span with reserved token ids
    ids[torch.arange(n), start] = 5
    ids[torch.arange(n), end] = 6
    return {"input_ids": ids, "token_type_ids": types, "attention_mask": mask, "start_positions": start, "end_positions": end}


feats = torch.load(args.features) if args.features else synthetic_features(args.num_synthetic, args.max_seq_length, cfg.vocab_size)
n = feats["input_ids"].shape[0]
n_eval = max(world * args.batch_size, n // 10)
keys = ["input_ids", "token_type_ids", "attention_mask", "start_positions", "end_positions"]
train_set = torch.utils.data.TensorDataset(*[feats[k][: n - n_eval] for k in keys])
eval_set = torch.utils.data.TensorDataset(*[feats[k][n - n_eval:] for k in keys])
if args.load_balance:
    sampler = LoadBalancingDistributedSampler(train_set, complexity_fn=lambda item: int(item[2].sum()), num_replicas=world, rank=rank, shuffle=True)
else:
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, num_replicas=world, rank=rank, shuffle=True)
train_loader = torch.utils.data.DataLoader(train_set, batch_size=args.batch_size, sampler=sampler, drop_last=True, pin_memory=cuda)
eval_loader = torch.utils.data.DataLoader(eval_set, batch_size=args.batch_size, drop_last=True, pin_memory=cuda,
                                          sampler=torch.utils.data.distributed.DistributedSampler(eval_set, num_replicas=world, rank=rank, shuffle=False))

model = models.BertForQuestionAnswering(cfg).to(dev).to(dtype)
if args.init:
    missing = model.load_state_dict(torch.load(args.init, map_location=dev), strict=False)
    if rank == 0:
        print("loaded", args.init, "missing:", len(missing.missing_keys), "unexpected:", len(missing.unexpected_keys))
decay = [p_ for n_, p_ in model.named_parameters() if p_.ndim > 1]
no_decay = [p_ for n_, p_ in model.named_parameters() if p_.ndim <= 1]
if args.algorithm == "qadam":
    optimizer = q_adam.QAdamOptimizer(model.parameters(), lr=args.learning_rate, warmup_steps=100)
    algorithm = q_adam.QAdamAlgorithm(optimizer)
elif args.fused_shard and cuda and world > 1:
    # AdamW inside the bucket kernels; the algorithm cuts the buckets so that each kernel serves one of the two groups
    optimizer = make_sharded_fused_adam([{"params": decay, "weight_decay": args.weight_decay}, {"params": no_decay, "weight_decay": 0.0}],
                                        lr=args.learning_rate, adamw=True)
    algorithm = FusedGradientAllReduceAlgorithm(optimizer)
else:
    groups = [{"params": decay, "weight_decay": args.weight_decay}, {"params": no_decay, "weight_decay": 0.0}]
    if cuda and not args.fuse_optimizer:
        optimizer = bagua.ops.FusedAdam(groups, lr=args.learning_rate, adamw=True)
    else:
        optimizer = torch.optim.AdamW(groups, lr=args.learning_rate)
    algorithm = Algorithm.init(args.algorithm)
model = model.with_bagua([optimizer], algorithm)
if args.fuse_optimizer:
    optimizer = bagua.contrib.fuse_optimizer(optimizer)
total_steps = max(1, args.epochs * len(train_loader))
warm = max(1, int(args.warmup_ratio * total_steps))
scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: (s + 1) / warm if s < warm else max(0.0, (total_steps - s) / max(1, total_steps - warm)))


def batch_to_device(batch):
    return [t.to(dev, non_blocking=True) for t in batch]


@torch.no_grad()
def evaluate():
    model.eval()
    stat = torch.zeros(2, device=dev)
    for batch in eval_loader:
        ids, tt, am, sp, ep = batch_to_device(batch)
        out = model(ids, token_type_ids=tt, attention_mask=am)
        start_logits, end_logits = out[-2], out[-1]
        ok = (start_logits.argmax(-1) == sp) & (end_logits.argmax(-1) == ep)
        stat += torch.stack([ok.float().sum(), torch.tensor(float(ok.numel()), device=dev)])
    bagua.allreduce_inplace(stat, op=bagua.ReduceOp.SUM)
    model.train()
    return (stat[0] / stat[1].clamp(min=1)).item()


step = 0
for epoch in range(args.epochs):
    sampler.set_epoch(epoch)
    if args.algorithm == "async":
        model.bagua_algorithm.resume(model)
    t0 = time.time()
    for batch in train_loader:
        ids, tt, am, sp, ep = batch_to_device(batch)
        optimizer.zero_grad()
        loss = model(ids, token_type_ids=tt, attention_mask=am, start_positions=sp, end_positions=ep)[0]
        loss.backward()
        optimizer.fuse_step() if args.fuse_optimizer else optimizer.step()
        scheduler.step()
        step += 1
        if step % args.print_freq == 0 and rank == 0:
            print(f"epoch {epoch} step {step}/{total_steps} loss {loss.item():.4f} lr {scheduler.get_last_lr()[0]:.3g} "
                  f"{args.batch_size * world * args.print_freq / (time.time() - t0):.1f} samples/s", flush=True)
            t0 = time.time()
    if args.algorithm == "async":
        model.bagua_algorithm.abort(model)
    em = evaluate()
    if rank == 0:
        print(f"epoch {epoch}: exact match {em * 100:.2f}", flush=True)
        if args.save_dir:
            os.makedirs(args.save_dir, exist_ok=True)
            torch.save(model.state_dict(), os.path.join(args.save_dir, f"bert_qa_epoch{epoch}.pt"))
