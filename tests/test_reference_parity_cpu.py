"""More of the reference's test scenarios on CPU/gloo: QAdam warm-up math (tests/torch_api/test_qadam.py),
find_unused_parameters (tests/torch_api/data_parallel/test_bagua_ddp.py), async abort/resume/abort
(tests/torch_api/test_async_model_average.py:77-86), optimizer-state broadcast for several optimizers
(tests/torch_api/test_broadcast_state.py), sub-group training (tests/torch_api/test_decentralized.py:406-422),
process-group bookkeeping (tests/torch_api/test_process_group.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.mp_utils import run_distributed


def test_qadam_optimizer_equals_adam_during_warmup():
    from bagua_b200.parallel.algorithms.q_adam import QAdamOptimizer

    torch.manual_seed(0)
    a = nn.Sequential(nn.Linear(10, 16), nn.Tanh(), nn.Linear(16, 3))
    b = nn.Sequential(nn.Linear(10, 16), nn.Tanh(), nn.Linear(16, 3))
    b.load_state_dict(a.state_dict())
    oa = QAdamOptimizer(a.parameters(), lr=1e-2, warmup_steps=50, weight_decay=0.01)
    ob = torch.optim.Adam(b.parameters(), lr=1e-2, weight_decay=0.01)
    for it in range(20):
        x = torch.randn(8, 10, generator=torch.Generator().manual_seed(it))
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            o.step()
    for p, q in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6)


class _Branchy(nn.Module):
    def __init__(self):
        super().__init__()
        self.trunk = nn.Linear(8, 8)
        self.a = nn.Linear(8, 4)
        self.b = nn.Linear(8, 4)      # only used when use_b
        self.never = nn.Linear(8, 4)  # never used

    def forward(self, x, use_b=False):
        h = torch.relu(self.trunk(x))
        return self.a(h) + (self.b(h) if use_b else 0)


def _unused_worker(rank, world):
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import gradient_allreduce
    from bagua_b200.parallel.data_parallel import DistributedDataParallel

    bagua.init_process_group()
    torch.manual_seed(3)
    model = _Branchy()
    ref = _Branchy()
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    ddp = DistributedDataParallel(model, optimizers=[opt], algorithm=gradient_allreduce.GradientAllReduceAlgorithm(), find_unused_parameters=True)
    for it in range(6):
        use_b = it >= 3                                       # the set of parameters in the graph changes mid-training
        x = torch.randn(4, 8, generator=torch.Generator().manual_seed(10 * it + rank))
        opt.zero_grad()
        ddp(x, use_b=use_b).pow(2).mean().backward()
        opt.step()
        ropt.zero_grad()
        ref(x, use_b=use_b).pow(2).mean().backward()
        for p in ref.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad)
                p.grad /= world
        ropt.step()
    mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    want = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    return mine, want


def test_find_unused_parameters_rebuckets_and_matches_manual_averaging():
    res = run_distributed(_unused_worker, world=2)
    for mine, want in res:
        torch.testing.assert_close(mine, want, rtol=1e-5, atol=1e-6)
    assert torch.equal(res[0][0], res[1][0])


def _async_worker(rank, world):
    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import async_model_average

    bagua.init_process_group()
    torch.manual_seed(rank)
    model = nn.Sequential(nn.Linear(6, 12), nn.ReLU(), nn.Linear(12, 2))
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    algo = async_model_average.AsyncModelAverageAlgorithm(sync_interval_ms=5, warmup_steps=2)
    model = model.with_bagua([opt], algo)
    impl = model.bagua_algorithm

    def epoch(n, seed):
        for it in range(n):
            x = torch.randn(4, 6, generator=torch.Generator().manual_seed(seed + it + 100 * rank))
            opt.zero_grad()
            model(x).pow(2).mean().backward()
            opt.step()

    for cycle in range(3):                                   # train / abort / (evaluate) / resume, repeatedly
        epoch(8, 1000 * cycle)
        impl.abort(model)
        impl.abort(model)                                     # a second abort is a no-op, not a hang
        with torch.no_grad():
            model(torch.zeros(1, 6))
        impl.resume(model)
        impl.resume(model)
    epoch(5, 9999)
    impl.abort(model)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    return flat, bool(torch.isfinite(flat).all())


def test_async_model_average_repeated_abort_resume():
    res = run_distributed(_async_worker, world=2, timeout=240)
    assert all(ok for _, ok in res)
    # the replicas have been averaged many times: they must be close (not bit-equal — the algorithm is asynchronous)
    assert (res[0][0] - res[1][0]).abs().max() < 0.5


def _broadcast_state_worker(rank, world):
    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import gradient_allreduce

    bagua.init_process_group()
    out = {}
    for name, make in [
        ("sgd_momentum", lambda ps: torch.optim.SGD(ps, lr=0.1 * (rank + 1), momentum=0.9)),
        ("adam", lambda ps: torch.optim.Adam(ps, lr=1e-3 * (rank + 1))),
        ("adamw", lambda ps: torch.optim.AdamW(ps, lr=1e-3 * (rank + 1), weight_decay=0.1 * (rank + 1))),
        ("rmsprop", lambda ps: torch.optim.RMSprop(ps, lr=1e-3 * (rank + 1), momentum=0.5)),
        ("adagrad", lambda ps: torch.optim.Adagrad(ps, lr=1e-2 * (rank + 1))),
    ]:
        torch.manual_seed(100 + rank)
        model = nn.Sequential(nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 3))
        opt = make(model.parameters())
        # give every rank a different, non-empty optimizer state before wrapping
        for it in range(2):
            opt.zero_grad()
            model(torch.randn(3, 5)).sum().backward()
            opt.step()
        model = model.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
        state = []
        for g in opt.param_groups:
            state.append(torch.tensor([float(v) for k, v in sorted(g.items()) if isinstance(v, (int, float)) and not isinstance(v, bool)]))
            for p in g["params"]:
                for k, v in sorted(opt.state[p].items()):
                    if torch.is_tensor(v):
                        state.append(v.detach().reshape(-1).float())
        out[name] = (torch.cat([p.detach().reshape(-1) for p in model.parameters()]), torch.cat(state))
    return out


def test_parameters_and_optimizer_state_are_broadcast_for_several_optimizers():
    res = run_distributed(_broadcast_state_worker, world=2, timeout=240)
    for name in res[0]:
        torch.testing.assert_close(res[0][name][0], res[1][name][0], rtol=0, atol=0, msg=name)
        torch.testing.assert_close(res[0][name][1], res[1][name][1], rtol=0, atol=0, msg=name)


def _subgroup_worker(rank, world):
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import decentralized, gradient_allreduce

    bagua.init_process_group()
    group = bagua.new_group(ranks=[0, 1])                     # every rank calls new_group; rank 2 is not a member
    tg = dist.new_group(ranks=[0, 1])
    if rank == 2:
        try:
            bagua.from_torch_group(tg)
            raise AssertionError("non-member must be rejected")
        except ValueError:
            return None
    assert bagua.from_torch_group(tg).ranks == [0, 1]
    torch.manual_seed(rank)
    outs = []
    for algo in (gradient_allreduce.GradientAllReduceAlgorithm(), decentralized.DecentralizedAlgorithm(hierarchical=False)):
        model = nn.Sequential(nn.Linear(4, 4), nn.ReLU(), nn.Linear(4, 2))
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        model = model.with_bagua([opt], algo, process_group=group)
        for it in range(4):
            x = torch.randn(4, 4, generator=torch.Generator().manual_seed(it + 7 * rank))
            opt.zero_grad()
            model(x).pow(2).mean().backward()
            opt.step()
        outs.append(torch.cat([p.detach().reshape(-1) for p in model.parameters()]))
    return outs


def test_training_on_a_sub_group_leaves_outsiders_alone():
    res = run_distributed(_subgroup_worker, world=3, timeout=240)
    assert res[2] is None
    torch.testing.assert_close(res[0][0], res[1][0], rtol=0, atol=0)          # gradient allreduce: bit-identical replicas
    assert (res[0][1] - res[1][1]).abs().max() < 0.2                          # decentralized: weights pulled together


def _pg_worker(rank, world):
    import gc

    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200 import communication as comm

    bagua.init_process_group()
    g = bagua.new_group(ranks=[0, 1])
    t = torch.full((4,), float(rank + 1))
    bagua.allreduce_inplace(t, comm=g.get_global_communicator())
    tg = dist.new_group(ranks=[0, 1])
    before = len(comm._torch_to_bagua_pg) if hasattr(comm, "_torch_to_bagua_pg") else None
    bg = bagua.from_torch_group(tg)
    same = bagua.from_torch_group(tg) is bg                                    # cached per torch group
    name = bg.group_name
    del bg
    gc.collect()
    return t, same, name, before


def test_process_group_bookkeeping():
    res = run_distributed(_pg_worker, world=2)
    for t, same, name, _ in res:
        assert torch.equal(t, torch.full((4,), 3.0)) and same and isinstance(name, str)


def _fused_combo_worker(rank, world, name):
    """fuse_optimizer together with an algorithm must train exactly like the unfused optimizer with the same algorithm; it
    fuses when the optimizer owns the flat layout (bagua do_flatten=False) and silently stays unfused when the bucket arena
    re-flattens the gradients in its own order (reference tests/contrib/test_fused_optimizer.py:317-437: 102 vs 0 fused steps)."""
    import bagua_b200 as bagua
    from bagua_b200.contrib import fuse_optimizer
    from bagua_b200.parallel.algorithms import Algorithm

    bagua.init_process_group()
    outs, counts = [], []
    for fused, bagua_flatten in ((False, True), (True, False), (True, True)):
        torch.manual_seed(5)
        model = nn.Sequential(nn.Linear(12, 24), nn.ReLU(), nn.Linear(24, 24), nn.ReLU(), nn.Linear(24, 4))
        opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
        if fused:
            opt = fuse_optimizer(opt, do_flatten=True)          # first fuse the optimizer, then wrap the module
        model = model.with_bagua([opt], Algorithm.init(name), do_flatten=bagua_flatten)
        for it in range(5):
            x = torch.randn(8, 12, generator=torch.Generator().manual_seed(100 * it + rank))
            opt.zero_grad()
            model(x).pow(2).mean().backward()
            opt.fuse_step() if fused else opt.step()
        outs.append(torch.cat([p.detach().reshape(-1) for p in model.parameters()]))
        counts.append(getattr(opt, "_bagua_fused_count", None))
    return outs, counts


import pytest  # noqa: E402


@pytest.mark.parametrize("name", ["gradient_allreduce", "bytegrad", "decentralized"])
def test_fuse_optimizer_combined_with_algorithms(name):
    res = run_distributed(_fused_combo_worker, world=2, args=(name,), timeout=240)
    for (plain, fused_optflat, fused_bothflat), counts in res:
        torch.testing.assert_close(plain, fused_optflat, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(plain, fused_bothflat, rtol=1e-6, atol=1e-7)
        assert counts[1] >= 1, counts      # optimizer-owned flat layout → the steps were fused


class _Tied(nn.Module):
    """Tied weights, a frozen layer, an ignored layer and a BatchNorm with buffers — the DDP corner cases of the reference's
    lifted PyTorch suite (tests/torch_api/data_parallel/test_c10d_common.py)."""

    def __init__(self):
        super().__init__()
        self.emb = nn.Linear(6, 6, bias=False)
        self.frozen = nn.Linear(6, 6)
        self.local_only = nn.Linear(6, 6)
        self.bn = nn.BatchNorm1d(6)
        self.out = nn.Linear(6, 6, bias=False)
        self.out.weight = self.emb.weight           # shared parameter: must be registered once
        for p in self.frozen.parameters():
            p.requires_grad_(False)
        self._bagua_params_and_buffers_to_ignore = ["local_only.weight", "local_only.bias"]

    def forward(self, x):
        return self.out(self.bn(self.local_only(self.frozen(self.emb(x)))))


def _ddp_corner_worker(rank, world):
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import gradient_allreduce

    bagua.init_process_group()
    torch.manual_seed(100 + rank)                    # different init (and BN buffers) per rank
    model = _Tied()
    with torch.no_grad():
        model.bn.running_mean.fill_(float(rank + 1))
    ref = _Tied()
    ref.load_state_dict(model.state_dict())
    for t in list(ref.parameters()) + list(ref.buffers()):
        if t.dtype.is_floating_point:
            dist.broadcast(t.data, 0)                # what with_bagua is expected to do ... except for the ignored layer
    with torch.no_grad():
        ref.local_only.load_state_dict(model.local_only.state_dict())
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.1)
    ropt = torch.optim.SGD([p for p in ref.parameters() if p.requires_grad], lr=0.1)
    model = model.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
    names = sorted(t.bagua_tensor_name for b in model.bagua_buckets for t in b.tensors if not t.bagua_tensor_name.startswith("bagua_padding"))
    for it in range(3):
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(50 * it + rank))
        for m, o, manual in ((model, opt, False), (ref, ropt, True)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            if manual:
                for n, p in ref.named_parameters():
                    if p.grad is not None and not n.startswith("local_only"):
                        dist.all_reduce(p.grad)
                        p.grad /= world
            o.step()
    mine = {n: p.detach().clone() for n, p in model.named_parameters()}
    want = {n: p.detach().clone() for n, p in ref.named_parameters()}
    return names, mine, want, model.bn.running_mean.clone()


def test_ddp_corner_cases_shared_frozen_ignored_parameters_and_buffers():
    res = run_distributed(_ddp_corner_worker, world=2, timeout=240)
    for names, mine, want, _ in res:
        assert names == ["bn.bias", "bn.weight", "emb.weight"], names   # tied weight once; frozen + ignored layers not registered
        for n in mine:
            torch.testing.assert_close(mine[n], want[n], rtol=1e-5, atol=1e-6, msg=n)
    # ignored parameters stay rank-local, everything else is identical on both ranks
    assert not torch.equal(res[0][1]["local_only.weight"], res[1][1]["local_only.weight"])
    for n in ("emb.weight", "bn.weight", "frozen.weight"):
        assert torch.equal(res[0][1][n], res[1][1][n]), n


def test_report_metrics_switch_logs_per_bucket_profile(tmp_path):
    """``BAGUA_REPORT_METRICS=1`` (launcher flag ``--report_metrics``) has a consumer here: periodic per-bucket communication report."""
    import os
    import subprocess
    import sys
    import textwrap

    from tests.mp_utils import free_port

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "m.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {repo!r})
        import torch, bagua_b200 as bagua
        from bagua_b200.parallel.algorithms import gradient_allreduce
        bagua.init_process_group()
        m = torch.nn.Linear(8, 8)
        opt = torch.optim.SGD(m.parameters(), lr=0.1)
        m = m.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
        m.bagua_ddp._report_every = 4
        for i in range(9):
            opt.zero_grad(); m(torch.randn(2, 8)).sum().backward(); opt.step()
    """))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), BAGUA_FORCE_CPU="1",
               CUDA_VISIBLE_DEVICES="", BAGUA_REPORT_METRICS="1")
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in (r.stdout + r.stderr).splitlines() if "[bagua metrics]" in ln]
    assert len(lines) == 2 and "step=4" in lines[0] and "step=8" in lines[1] and "'launches': 4" in lines[1], lines


def _zero_grad_worker(rank, world):
    """``model.zero_grad()`` and ``optimizer.zero_grad()`` (both default to set_to_none=True in current torch) keep the gradients
    allocated as bucket views, so training continues without tripping the pointer contract."""
    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import gradient_allreduce

    bagua.init_process_group()
    torch.manual_seed(3)
    model = torch.nn.Linear(6, 3)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    model = model.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
    ptrs = None
    for i in range(4):
        (model.zero_grad if i % 2 else opt.zero_grad)()
        assert all(p.grad is not None and float(p.grad.abs().sum()) == 0.0 for p in model.parameters())
        model(torch.randn(5, 6)).sum().backward()
        opt.step()
        now = [p.grad.data_ptr() for p in model.parameters()]
        assert ptrs is None or now == ptrs
        ptrs = now
    # the DDP-compatible wrapper: its own nn.Module.zero_grad must not drop the views either
    from bagua_b200.parallel.data_parallel import DistributedDataParallel

    inner = torch.nn.Linear(6, 3)
    opt2 = torch.optim.SGD(inner.parameters(), lr=0.1)
    ddp = DistributedDataParallel(inner, optimizers=[opt2])
    for _ in range(3):
        ddp.zero_grad()
        ddp(torch.randn(5, 6)).sum().backward()
        opt2.step()
        assert all(p.grad is not None for p in inner.parameters())
    return True


def test_zero_grad_keeps_bucket_views():
    assert all(run_distributed(_zero_grad_worker, world=2))
