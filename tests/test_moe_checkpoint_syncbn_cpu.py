"""MoE gating/dispatch (index form vs the reference's dense one-hot formulation), expert-parallel training on gloo,
MoE-aware checkpoints, and SyncBatchNorm against single-process BatchNorm over the concatenated batch."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from tests.mp_utils import run_distributed


def _dense_reference(tokens, gate, E, C):
    """The reference's formulation (sharded_moe.py:352-374) from dense combine weights."""
    cw, mask = gate.dense()
    dispatched = torch.einsum("sec,sm->ecm", mask.type_as(tokens), tokens)
    return cw, mask, dispatched


@pytest.mark.parametrize("k", [1, 2])
def test_index_gating_matches_dense_formulation(k):
    from bagua_b200.ops import moe as moe_ops
    from bagua_b200.parallel.moe.sharded_moe import top1gating_indices, top2gating_indices

    torch.manual_seed(0)
    S, E, M = 64, 4, 8
    logits = torch.randn(S, E)
    tokens = torch.randn(S, M, requires_grad=True)
    g = top1gating_indices(logits, 1.0, 4) if k == 1 else top2gating_indices(logits, 1.0)
    C = g.capacity
    assert C == (math.ceil(S / E) if k == 1 else math.ceil(2 * S / E))
    # every (expert, slot) pair is used at most once and slots are < capacity
    flat = (g.expert_idx * C + g.slot_idx)[g.slot_idx >= 0]
    assert flat.unique().numel() == flat.numel() and int(g.slot_idx.max()) < C
    cw, mask, dispatched_ref = _dense_reference(tokens.detach(), g, E, C)
    disp = moe_ops.dispatch(tokens, g.expert_idx, g.slot_idx, E, C, None, 1, E).reshape(E, C, M)
    torch.testing.assert_close(disp, dispatched_ref)
    expert_out = torch.randn(1, E, C, M, requires_grad=True)
    w = g.weights.clone().requires_grad_(True)
    comb = moe_ops.combine(expert_out, g.expert_idx, g.slot_idx, w, E, C, None, 1, E)
    comb_ref = torch.einsum("sec,ecm->sm", cw, expert_out.detach().reshape(E, C, M))
    torch.testing.assert_close(comb, comb_ref, rtol=1e-5, atol=1e-6)
    # gradients of the row scatter/gather equal autograd through the dense einsums
    (disp.sum() + comb.pow(2).sum()).backward()
    t2 = tokens.detach().clone().requires_grad_(True)
    e2 = expert_out.detach().clone().requires_grad_(True)
    cw2 = cw.clone().requires_grad_(True)
    ref = torch.einsum("sec,sm->ecm", mask.float(), t2).sum() + torch.einsum("sec,ecm->sm", cw2, e2.reshape(E, C, M)).pow(2).sum()
    ref.backward()
    torch.testing.assert_close(tokens.grad, t2.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(expert_out.grad, e2.grad, rtol=1e-5, atol=1e-6)
    gw_ref = torch.stack([cw2.grad[torch.arange(S), g.expert_idx[:, j], g.slot_idx[:, j].clamp(min=0)] * (g.slot_idx[:, j] >= 0) for j in range(k)], dim=1)
    torch.testing.assert_close(w.grad, gw_ref, rtol=1e-5, atol=1e-6)
    # l_aux and exp_counts as defined by the reference
    gates = F.softmax(logits, dim=1)
    mask1 = F.one_hot(gates.argmax(1), E)
    assert torch.equal(g.exp_counts, mask1.sum(0))
    l_ref = (gates.mean(0) * mask1.float().mean(0)).sum() * E if k == 1 else (gates.mean(0) * mask1.float().mean(0)).mean() * E * E
    torch.testing.assert_close(g.l_aux, l_ref)


def test_top1_capacity_overflow_drops_tokens():
    from bagua_b200.parallel.moe.sharded_moe import top1gating_indices

    logits = torch.zeros(32, 4)
    logits[:, 0] = 10.0  # everybody wants expert 0
    g = top1gating_indices(logits, 1.0, 4)
    kept = (g.slot_idx >= 0).sum().item()
    assert kept == g.capacity == 8 and (g.weights[g.slot_idx < 0] == 0).all()


def _moe_worker(rank, world, ckpt_dir):
    import torch.distributed as dist
    import torch.nn as nn

    import bagua_b200 as bagua
    from bagua_b200.checkpoint import load_checkpoint, save_checkpoint
    from bagua_b200.parallel.algorithms import gradient_allreduce
    from bagua_b200.parallel.moe import MoE

    bagua.init_process_group()

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.inp = nn.Linear(16, 32)
            self.moe = MoE(32, nn.Linear(32, 32), num_local_experts=2, k=2)
            self.out = nn.Linear(32, 4)

        def forward(self, x):
            h, l_aux, counts = self.moe(F.relu(self.inp(x)))
            return self.out(h), l_aux, counts

    torch.manual_seed(100 + rank)  # experts start different per rank; shared weights are broadcast by with_bagua
    model = Net()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    model = model.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
    names = {n for n, _ in model.bagua_ddp.bagua_build_params()}
    assert not any("bagua_experts" in n for n in names) and "moe.bagua_moe.gate.wg.weight" in names
    for it in range(4):
        g = torch.Generator().manual_seed(rank * 10 + it)
        x, y = torch.randn(16, 16, generator=g), torch.randint(0, 4, (16,), generator=g)
        opt.zero_grad()
        out, l_aux, counts = model(x)
        (F.cross_entropy(out, y) + 0.01 * l_aux).backward()
        opt.step()
        assert counts.device.type == "cpu" and counts.numel() == 2 * world
    shared = torch.cat([p.detach().reshape(-1) for n, p in model.named_parameters() if "bagua_experts" not in n])
    expert = torch.cat([p.detach().reshape(-1) for n, p in model.named_parameters() if "bagua_experts" in n])
    save_checkpoint(7, ckpt_dir, model, opt)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for p in model.parameters():
            p.add_(1.0)
    it = load_checkpoint(ckpt_dir, model, opt)
    after = model.state_dict()
    assert it == 7 and all(torch.equal(before[k], after[k]) for k in before)
    files = sorted(os.listdir(os.path.join(ckpt_dir, "iter_0000007"))) if rank == 0 else []
    return shared, expert, files


def test_moe_training_and_checkpoint(tmp_path):
    res = run_distributed(_moe_worker, world=2, args=(str(tmp_path),), timeout=300)
    assert torch.equal(res[0][0], res[1][0])          # non-expert parameters are data-parallel → identical
    assert not torch.equal(res[0][1], res[1][1])      # experts are local to their rank
    files = res[0][2]
    assert "mp_rank_00_model_states.pt" in files
    assert {f"expert_{i}_mp_rank_00_model_states.pt" for i in range(4)} <= set(files)
    assert {f"expert_parallel_rank_{r}_mp_rank_00_optim_states.pt" for r in range(2)} <= set(files)
    assert open(tmp_path / "latest_checkpointed_iteration.txt").read().strip() == "7"


def test_plain_checkpoint_roundtrip(tmp_path):
    from bagua_b200.checkpoint import load_checkpoint, save_checkpoint

    m = torch.nn.Linear(4, 4)
    opt = torch.optim.Adam(m.parameters())
    m(torch.randn(2, 4)).sum().backward()
    opt.step()
    sched = torch.optim.lr_scheduler.StepLR(opt, 1)
    assert load_checkpoint(str(tmp_path), m) == 0  # nothing there yet
    save_checkpoint(3, str(tmp_path), m, opt, sched)
    m2 = torch.nn.Linear(4, 4)
    opt2 = torch.optim.Adam(m2.parameters())
    assert load_checkpoint(str(tmp_path), m2, opt2, torch.optim.lr_scheduler.StepLR(opt2, 1)) == 3
    assert torch.equal(m.weight, m2.weight) and len(opt2.state_dict()["state"]) == 2


def _syncbn_worker(rank, world):
    import bagua_b200 as bagua
    from bagua_b200.contrib.sync_batchnorm import SyncBatchNorm

    bagua.init_process_group()
    torch.manual_seed(0)
    full = torch.randn(4 * world, 3, 5, 5)
    ref_bn = torch.nn.BatchNorm2d(3)
    net = torch.nn.Sequential(torch.nn.BatchNorm2d(3))
    net = SyncBatchNorm.convert_sync_batchnorm(net)
    assert isinstance(net[0], SyncBatchNorm)
    x = full[4 * rank : 4 * rank + 4].clone().requires_grad_(True)
    xf = full.clone().requires_grad_(True)
    y = net(x)
    yf = ref_bn(xf)
    w = torch.linspace(0.5, 1.5, y.numel() * world).view(4 * world, 3, 5, 5)
    (y * w[4 * rank : 4 * rank + 4]).sum().backward()
    (yf * w).sum().backward()
    torch.testing.assert_close(y, yf[4 * rank : 4 * rank + 4], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x.grad, xf.grad[4 * rank : 4 * rank + 4], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(net[0].running_mean, ref_bn.running_mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(net[0].running_var, ref_bn.running_var, rtol=1e-4, atol=1e-5)
    net.eval()
    assert torch.isfinite(net(x.detach())).all()
    return True


def test_sync_batchnorm_equals_global_batchnorm():
    assert all(run_distributed(_syncbn_worker, world=2))


def _bf16_moe_worker(rank, world):
    """Regression: a bf16 model whose MoE gate must stay fp32 trains under with_bagua (casting the gate inside ``forward`` used to
    re-create the parameter and break the engine's 'backend tensor == parameter grad' contract)."""
    import bagua_b200 as bagua
    from bagua_b200 import models
    from bagua_b200.parallel.algorithms import gradient_allreduce

    bagua.init_process_group()
    torch.manual_seed(0)
    cfg = models.GPT2MoEConfig(vocab_size=128, n_positions=16, n_embd=32, n_layer=2, n_head=2, num_experts=2 * world)
    model = models.GPT2MoE(cfg, world_size=world).to(torch.bfloat16)
    gate_dtypes = {p.dtype for n, p in model.named_parameters() if n.endswith("wg.weight")}
    assert gate_dtypes == {torch.float32}
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    model = model.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
    idx = torch.randint(0, 128, (2, 16), generator=torch.Generator().manual_seed(rank))
    for _ in range(3):
        opt.zero_grad()
        loss, _ = model(idx, idx)
        loss.backward()
        opt.step()
    gate = [p for n, p in model.named_parameters() if n.endswith("wg.weight")][0]
    return float(loss), gate.detach().clone()


def test_bf16_model_with_fp32_gate_trains_under_with_bagua():
    res = run_distributed(_bf16_moe_worker, world=2, timeout=240)
    assert all(l == l for l, _ in res)                                  # finite
    assert torch.equal(res[0][1], res[1][1])                            # the gate is data-parallel: identical on both ranks
