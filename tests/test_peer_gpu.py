"""Multi-GPU tests of the NVSwitch peer kernels (≥ 2 GPUs): every fused collective against torch.distributed / the
python oracles, then the algorithms end to end."""
import pytest
import torch

from tests.mp_utils import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ngpu():
    return min(torch.cuda.device_count(), 8)


def _allreduce_worker(rank, world):
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.core import native

    bagua.init_process_group()
    pg = bagua.communication._get_default_group()
    eng = pg.peer_engine()
    assert eng is not None, "peer engine must be available on a single NVSwitch node"
    dev = torch.device("cuda", rank)
    C = native()
    results = {"multicast": eng.has_multicast}
    stream = torch.cuda.current_stream().cuda_stream
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        for numel in (4096, 1 << 20, (1 << 22) + 4096):
            nbytes = numel * torch.empty(0, dtype=dtype).element_size()
            sl = eng.alloc(nbytes)
            for variant in ["one_shot", "two_shot"] + (["multimem"] if eng.has_multicast else []):
                torch.manual_seed(100 + rank)
                x = torch.randn(numel, device=dev).to(dtype)
                t = sl.view(dtype, numel)
                t.copy_(x)
                ref = x.float().clone()
                dist.all_reduce(ref)
                ref /= world
                op, chosen = eng.make_allreduce_op(sl, sl, nbytes, dtype, True, variant)
                torch.cuda.synchronize()
                dist.barrier()
                C.run_op(op, stream, rank)
                torch.cuda.synchronize()
                tol = 1e-5 if dtype == torch.float32 else 2e-2
                err = (t.float() - ref).abs().max().item()
                assert err <= tol * max(1.0, ref.abs().max().item()), f"{dtype} {numel} {variant}->{chosen}: err {err}"
            sl.free()
    # blocking API fast path on an arbitrary tensor
    y = torch.full((12345,), float(rank + 1), device=dev)
    bagua.allreduce_inplace(y, op=bagua.ReduceOp.AVG)
    assert torch.allclose(y, torch.full_like(y, (world + 1) / 2))
    assert eng.comm.error_code() == 0
    return results


def test_peer_allreduce_variants():
    run_distributed(_allreduce_worker, world=_ngpu(), use_cuda=True)


def _bytegrad_worker(rank, world):
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.core import dtype_code, native
    from bagua_b200.ops import quant

    bagua.init_process_group()
    pg = bagua.communication._get_default_group()
    eng = pg.peer_engine()
    C = native()
    dev = torch.device("cuda", rank)
    for dtype in (torch.float32, torch.bfloat16):
        numel = 32 * world * 257
        torch.manual_seed(7 + rank)
        x = torch.randn(numel, device=dev).to(dtype)
        # oracle: the reference pipeline on torch.distributed
        class FakeBucket:
            def __init__(self, t):
                self.t = t

            def _flat_or_gather(self):
                return self.t, None

        ref = x.clone()
        quant.bytegrad_allreduce_fallback(FakeBucket(ref), pg, True)
        data = x.clone()
        box = C.ByteGradOp.box_bytes(numel, world)
        inbox, outbox = eng.alloc(box), eng.alloc(box)
        op = C.ByteGradOp(eng.comm, data.data_ptr(), numel, dtype_code(dtype), inbox.buf, inbox.offset, outbox.buf, outbox.offset, True, C.LaunchCfg(2 * world, 256))
        torch.cuda.synchronize()
        dist.barrier()
        for _ in range(2):  # run twice: exercises the parity double-buffering
            data.copy_(x)
            C.run_op(op, torch.cuda.current_stream().cuda_stream, rank)
        torch.cuda.synchronize()
        step = (x.float().max() - x.float().min()).item() / 255
        assert (data.float() - ref.float()).abs().max().item() <= 2.5 * step
        assert eng.comm.error_code() == 0
    return True


def test_bytegrad_fused_matches_pipeline():
    run_distributed(_bytegrad_worker, world=_ngpu(), use_cuda=True)


def _algorithms_worker(rank, world, name):
    import torch.distributed as dist
    import torch.nn as nn

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import Algorithm, q_adam

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    torch.manual_seed(rank)
    model = nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 16)).to(dev)
    if name == "qadam":
        opt = q_adam.QAdamOptimizer(model.parameters(), lr=1e-3, warmup_steps=3)
        algo = q_adam.QAdamAlgorithm(opt)
    else:
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        kw = {"sync_interval_ms": 20} if name == "async" else {}
        if name == "decentralized_shift_one":
            algo = Algorithm.init("decentralized", peer_selection_mode="shift_one")
        else:
            algo = Algorithm.init(name, **kw)
    model = model.with_bagua([opt], algo)
    for it in range(8):
        x = torch.randn(8, 32, device=dev)
        opt.zero_grad()
        model(x).square().mean().backward()
        opt.step()
    if name == "async":
        model.bagua_algorithm.abort(model)
    torch.cuda.synchronize()
    flat = torch.cat([p.data.view(-1) for p in model.parameters()])
    assert torch.isfinite(flat).all()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    spread = max((g - gathered[0]).abs().max().item() for g in gathered)
    eng = bagua.communication._get_default_group().peer_engine()
    assert eng.comm.error_code() == 0
    return spread


@pytest.mark.parametrize("name", ["gradient_allreduce", "bytegrad", "decentralized", "decentralized_shift_one", "low_precision_decentralized", "qadam", "async"])
def test_algorithms_on_peer_kernels(name):
    world = _ngpu() if _ngpu() % 2 == 0 else _ngpu() - 1
    spreads = run_distributed(_algorithms_worker, world=world, args=(name,), use_cuda=True)
    if name == "gradient_allreduce":
        assert max(spreads) < 1e-6  # replicas stay bit-close
    elif name in ("bytegrad", "qadam"):
        assert max(spreads) < 1e-2


def _moe_worker(rank, world):
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.ops import moe as moe_ops
    from bagua_b200.ops import moe_peer
    from bagua_b200.parallel.moe.sharded_moe import top2gating_indices

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    torch.manual_seed(5 + rank)
    S, M, E_local = 256, 128, 2
    E = E_local * world
    for dtype in (torch.float32, torch.bfloat16):
        logits = torch.randn(S, E, device=dev)
        g = top2gating_indices(logits, 1.0)
        C = g.capacity
        tokens = torch.randn(S, M, device=dev).to(dtype)
        expert_w = torch.randn(M, M, device=dev).to(dtype) * 0.1
        outs = []
        for peer in (True, False):
            os_env = "1" if peer else "0"
            import os

            os.environ["BAGUA_MOE_PEER"] = os_env
            moe_peer._contexts.clear()
            t = tokens.clone().requires_grad_(True)
            w = g.weights.to(dtype).clone().requires_grad_(True)
            ew = expert_w.clone().requires_grad_(True)
            disp = moe_ops.dispatch(t, g.expert_idx, g.slot_idx, E, C, dist.group.WORLD, world, E_local)
            assert disp.shape == (world, E_local, C, M)
            eo = disp @ ew
            out = moe_ops.combine(eo, g.expert_idx, g.slot_idx, w, E, C, dist.group.WORLD, world, E_local)
            out.float().pow(2).sum().backward()
            outs.append((out.detach().float(), t.grad.float(), w.grad.float(), ew.grad.float()))
            if peer:
                assert moe_peer.get_context(dist.group.WORLD, world) is not None, "peer MoE path must be active"
        tol = 1e-4 if dtype == torch.float32 else 6e-2
        for a, b in zip(*outs):
            torch.testing.assert_close(a, b, rtol=tol, atol=tol * max(1.0, b.abs().max().item()))
    return True


def test_moe_peer_dispatch_combine_matches_all_to_all():
    run_distributed(_moe_worker, world=_ngpu(), use_cuda=True)


def _fused_combine_worker(rank, world):
    """fc2 GEMM with the combine all-to-all in its epilogue vs the two-step path (grouped GEMM, then pull-combine)."""
    import os

    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.models.gpt2_moe import ExpertMLP
    from bagua_b200.ops import moe as moe_ops
    from bagua_b200.ops import moe_peer
    from bagua_b200.parallel.moe.experts import Experts
    from bagua_b200.parallel.moe.sharded_moe import top2gating_indices

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    torch.manual_seed(11 + rank)
    S, M, E_local = 512, 256, 2
    E = E_local * world
    logits = torch.randn(S, E, device=dev)
    g = top2gating_indices(logits, 1.0)
    C = ((g.capacity + 127) // 128) * 128      # the 128-row GEMM tiles need a capacity that is a multiple of 128
    slot_idx = g.slot_idx
    experts = Experts(ExpertMLP(M), E_local).to(dev).to(torch.bfloat16)
    tokens = (torch.randn(S, M, device=dev) * 0.5).to(torch.bfloat16)
    res = []
    for fused in (True, False):
        os.environ["BAGUA_MOE_FUSED_COMBINE"] = "1" if fused else "0"
        experts.zero_grad(set_to_none=True)
        t = tokens.clone().requires_grad_(True)
        w = g.weights.to(torch.bfloat16).clone().requires_grad_(True)
        disp = moe_ops.dispatch(t, g.expert_idx, slot_idx, E, C, dist.group.WORLD, world, E_local)
        pctx = experts.fused_combine_context(disp, dist.group.WORLD, world)
        assert (pctx is not None) == fused
        if fused:
            out = experts.forward_combine(disp, w, g.expert_idx, slot_idx, pctx)
        else:
            out = moe_ops.combine(experts(disp), g.expert_idx, slot_idx, w, E, C, dist.group.WORLD, world, E_local)
        out.float().pow(2).sum().backward()
        res.append([out.detach().float(), t.grad.float(), w.grad.float()] + [p.grad.float() for p in experts.parameters()])
    for a, b in zip(*res):
        torch.testing.assert_close(a, b, rtol=5e-2, atol=5e-2 * max(1.0, b.abs().max().item()))
    assert moe_peer.get_context(dist.group.WORLD, world).comm.error_code() == 0
    return True


def test_moe_fused_gemm_combine_matches_two_step():
    run_distributed(_fused_combine_worker, world=_ngpu(), use_cuda=True)


def _fused_adam_worker(rank, world):
    """reduce-scatter → AdamW on the shard → all-gather in one kernel per bucket vs torch.optim.AdamW on averaged gradients."""
    import copy

    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms.gradient_allreduce import FusedGradientAllReduceAlgorithm, make_sharded_fused_adam

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    torch.manual_seed(7)
    model = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 512), torch.nn.GELU(), torch.nn.Linear(512, 64)).to(dev)
    oracle = copy.deepcopy(model)
    opt = make_sharded_fused_adam(model.parameters(), lr=1e-3, weight_decay=0.01, adamw=True)
    oopt = torch.optim.AdamW(oracle.parameters(), lr=1e-3, weight_decay=0.01)
    model = model.with_bagua([opt], FusedGradientAllReduceAlgorithm(opt))
    assert all(b.allreduce_variant.startswith("fused_adam") for b in model.bagua_buckets)
    for it in range(5):
        x = torch.randn(32, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + rank))
        opt.zero_grad()
        model(x).pow(2).mean().backward()
        opt.step()
        oopt.zero_grad()
        oracle(x).pow(2).mean().backward()
        for p in oracle.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        oopt.step()
    torch.cuda.synchronize()
    mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    want = torch.cat([p.detach().reshape(-1) for p in oracle.parameters()])
    torch.testing.assert_close(mine, want, rtol=1e-4, atol=1e-5)
    return mine


def test_fused_allreduce_adam_matches_adamw():
    res = run_distributed(_fused_adam_worker, world=_ngpu(), use_cuda=True)
    for r in res[1:]:
        assert torch.equal(res[0], r)


def _hier_worker(rank, world):
    """One box pretending to be 2 nodes of world/2 GPUs: NVLink reduce-scatter kernel → NCCL all-reduce on the rail → NVLink
    all-gather kernel must equal a flat all-reduce."""
    import os

    L = world // 2
    os.environ["NODE_RANK"] = str(rank // L)
    os.environ["LOCAL_WORLD_SIZE"] = str(L)
    import copy

    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import gradient_allreduce

    bagua.init_process_group()
    pg = bagua.communication._get_default_group()
    assert pg.nnodes == 2 and pg.peer_engine() is None and pg.hier_engine() is not None
    dev = torch.device("cuda", rank)
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(128, 256), torch.nn.ReLU(), torch.nn.Linear(256, 64)).to(dev)
    oracle = copy.deepcopy(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    oopt = torch.optim.SGD(oracle.parameters(), lr=0.05)
    model = model.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm(hierarchical=True))
    assert all(b.allreduce_variant.startswith("hier:") for b in model.bagua_buckets), [b.allreduce_variant for b in model.bagua_buckets]
    for it in range(4):
        x = torch.randn(16, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(10 * it + rank))
        for m, o, manual in ((model, opt, False), (oracle, oopt, True)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            if manual:
                for p in oracle.parameters():
                    dist.all_reduce(p.grad)
                    p.grad /= world
            o.step()
    torch.cuda.synchronize()
    mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    want = torch.cat([p.detach().reshape(-1) for p in oracle.parameters()])
    torch.testing.assert_close(mine, want, rtol=1e-5, atol=1e-6)
    return mine


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="hierarchical (virtual multi-node) path needs >= 4 GPUs posing as 2 nodes")
def test_hierarchical_allreduce_on_virtual_nodes():
    n = 4 if _ngpu() < 8 else 8
    res = run_distributed(_hier_worker, world=n, use_cuda=True)
    for r in res[1:]:
        assert torch.equal(res[0], r)


def _abort_worker(rank, world):
    """Fault injection (reference tests/comm/test_communicator.py:48-70): an all-reduce nobody else joins must be unblocked by
    ``abort()``, and a second one by the in-kernel timeout — the GPU is never left spinning."""
    import time

    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.core import native

    bagua.init_process_group()
    pg = bagua.communication._get_default_group()
    eng = pg.peer_engine()
    C = native()
    nbytes = 1 << 20
    sl = eng.alloc(nbytes)
    op, _ = eng.make_allreduce_op(sl, sl, nbytes, torch.float32, True, "two_shot")
    dist.barrier()
    out = {}
    if rank == 0:
        stream = torch.cuda.current_stream()
        C.run_op(op, stream.cuda_stream, rank)
        time.sleep(1.0)
        out["spinning"] = not stream.query()
        t0 = time.time()
        pg.get_global_communicator().abort()
        torch.cuda.synchronize()
        out["abort_latency_s"] = time.time() - t0
        out["code_after_abort"] = eng.comm.error_code()
        eng.comm.reset_abort()
        eng.comm.clear_error()
        eng.comm.set_timeout(2.0)
        t0 = time.time()
        C.run_op(op, stream.cuda_stream, rank)
        torch.cuda.synchronize()
        out["timeout_latency_s"] = time.time() - t0
        out["code_after_timeout"] = eng.comm.error_code()
        try:
            C.run_op(op, stream.cuda_stream, rank)
            out["fatal_after_timeout"] = False
        except RuntimeError as e:
            out["fatal_after_timeout"] = "timed out waiting for another rank" in str(e)
        eng.comm.clear_error()
    dist.barrier()
    return out


def test_abort_and_timeout_unblock_a_lonely_allreduce():
    res = run_distributed(_abort_worker, world=2, use_cuda=True)
    r0 = res[0]
    assert r0["spinning"] and r0["abort_latency_s"] < 1.0 and r0["code_after_abort"] == 2
    assert 1.5 < r0["timeout_latency_s"] < 6.0 and r0["code_after_timeout"] in (1, 3)
    assert r0["fatal_after_timeout"], "a timed-out collective must make every later op of the communicator raise"


def _syncbn_worker(rank, world):
    """CUDA path of contrib.SyncBatchNorm (ATen batch_norm_* kernels + packed all-gather / all-reduce) vs torch.nn.SyncBatchNorm
    (reference tests/contrib/test_sync_bn.py:66-227)."""
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.contrib.sync_batchnorm import SyncBatchNorm

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    torch.manual_seed(20 + rank)
    mine = SyncBatchNorm(16).to(dev)
    ref = torch.nn.SyncBatchNorm(16).to(dev)
    with torch.no_grad():
        w, b = torch.rand(16, device=dev) + 0.5, torch.randn(16, device=dev)
        dist.broadcast(w, 0)
        dist.broadcast(b, 0)
        for m in (mine, ref):
            m.weight.copy_(w)
            m.bias.copy_(b)
    outs = []
    for m in (mine, ref):
        x = torch.randn(4 + rank, 16, 5, 7, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)).requires_grad_(True)
        y = m(x)
        (y * torch.arange(y.numel(), device=dev).view_as(y).float().cos()).sum().backward()
        outs.append((y.detach(), x.grad, m.weight.grad, m.bias.grad, m.running_mean.clone(), m.running_var.clone()))
    for a, b_ in zip(*outs):
        torch.testing.assert_close(a, b_, rtol=1e-4, atol=1e-4)
    return True


def test_sync_batchnorm_cuda_matches_torch():
    run_distributed(_syncbn_worker, world=min(_ngpu(), 4), use_cuda=True)


def _peer_collectives_worker(rank, world):
    """Opt-in peer-kernel all-gather / reduce-scatter (``BAGUA_PEER_COLLECTIVES=1``) against torch.distributed."""
    import os

    import torch.distributed as dist

    os.environ["BAGUA_PEER_COLLECTIVES"] = "1"
    import bagua_b200 as bagua

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    for dtype in (torch.float32, torch.bfloat16, torch.int64):
        for n in (8, 4096, 1 << 20):
            g = torch.Generator(device=dev).manual_seed(1000 * rank + n)
            send = (torch.randn(n, device=dev, generator=g) * 100).to(dtype)
            got = torch.empty(n * world, dtype=dtype, device=dev)
            want = torch.empty_like(got)
            bagua.allgather(send, got)
            dist.all_gather_into_tensor(want, send)
            assert torch.equal(got, want), (dtype, n)
    for dtype in (torch.float32, torch.bfloat16):
        for n in (8, 4096, 1 << 18):
            g = torch.Generator(device=dev).manual_seed(7 * rank + n)
            send = torch.randn(n * world, device=dev, generator=g).to(dtype)
            got = torch.empty(n, dtype=dtype, device=dev)
            want = torch.empty(n, dtype=torch.float32, device=dev)
            bagua.reduce_scatter(send, got, op=bagua.ReduceOp.AVG)
            dist.reduce_scatter_tensor(want, send.float())
            want /= world
            tol = 1e-5 if dtype == torch.float32 else 2e-2
            torch.testing.assert_close(got.float(), want, rtol=tol, atol=tol * max(1.0, want.abs().max().item()))
    torch.cuda.synchronize()
    assert bagua.communication._get_default_group().peer_engine().comm.error_code() == 0
    return True


def test_peer_allgather_and_reduce_scatter_match_torch():
    run_distributed(_peer_collectives_worker, world=_ngpu(), use_cuda=True)


def _graphed_step_worker(rank, world):
    """Whole step (forward, backward, bucket all-reduce kernels, fused SGD) replayed from a CUDA graph on every rank, against eager
    twins: same losses and weights; the buckets were issued inline (no worker thread inside the capture)."""
    import copy

    import torch.nn.functional as F

    import bagua_b200 as bagua
    from bagua_b200.ops.optim import FusedSGD
    from bagua_b200.parallel.algorithms import gradient_allreduce
    from bagua_b200.utils.graph import GraphedTrainStep

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    torch.manual_seed(5)
    base = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.ReLU(), torch.nn.Linear(512, 512), torch.nn.ReLU(), torch.nn.Linear(512, 16)).to(dev)

    def make():
        m = copy.deepcopy(base)
        opt = FusedSGD(m.parameters(), lr=0.05, momentum=0.9)
        m = m.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())

        def step(x, y):
            opt.zero_grad()
            loss = F.cross_entropy(m(x), y)
            loss.backward()
            opt.step()
            return loss

        return m, opt, step

    torch.manual_seed(100 + rank)
    batches = [(torch.randn(32, 256, device=dev), torch.randint(0, 16, (32,), device=dev)) for _ in range(6)]
    m1, o1, eager = make()
    m2, o2, step2 = make()
    g = GraphedTrainStep(m2, step2, batches[0], optimizers=[o2], warmup=3)
    assert g.communicates
    for _ in range(3):
        eager(*batches[0])
    out = []
    for x, y in batches:
        le, lg = eager(x, y), g(x, y).clone()
        out.append((float(le), float(lg)))
    torch.cuda.synchronize()
    be = m2.bagua_ddp._bagua_backend
    assert be.inline_mode() and g.captures == 1 and g.replays == len(batches)
    for le, lg in out:
        assert abs(le - lg) <= 1e-4 * max(1.0, abs(le)), out
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-5)
    return out


def test_graphed_step_with_bucket_communication_matches_eager():
    res = run_distributed(_graphed_step_worker, world=_ngpu(), use_cuda=True)
    assert len(res) == _ngpu()


def _sharded_state_worker(rank, world):
    """Checkpoint / resume of the in-bucket SGD (momentum) and Adam: train, snapshot (model + consolidated optimizer state), train
    on, rewind to the snapshot, train the same steps again → identical weights; the consolidated state is the same on every rank."""
    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms.gradient_allreduce import FusedGradientAllReduceAlgorithm, make_sharded_fused_adam, make_sharded_fused_sgd

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    out = []
    for make in (lambda ps: make_sharded_fused_sgd(ps, lr=0.05, momentum=0.9), lambda ps: make_sharded_fused_adam(ps, lr=1e-3, adamw=True, weight_decay=0.01)):
        torch.manual_seed(21)
        model = torch.nn.Sequential(torch.nn.Linear(128, 256), torch.nn.ReLU(), torch.nn.Linear(256, 32)).to(dev)
        opt = make(model.parameters())
        model = model.with_bagua([opt], FusedGradientAllReduceAlgorithm(opt))

        def steps(first, n):
            for it in range(first, first + n):
                x = torch.randn(16, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(1000 * it + rank))
                opt.zero_grad()
                model(x).pow(2).mean().backward()
                opt.step()
            torch.cuda.synchronize()
            return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()

        steps(0, 3)
        assert opt.collective_state_dict
        snap_opt = opt.state_dict()
        snap_model = {k: v.clone() for k, v in model.state_dict().items()}
        assert snap_opt["steps"] == 3 and set(snap_opt["state"]) >= {"0.weight", "2.bias"}
        after = steps(3, 2)
        model.load_state_dict(snap_model)
        opt.load_state_dict(snap_opt)
        again = steps(3, 2)
        torch.testing.assert_close(again, after, rtol=1e-5, atol=1e-6)   # in-switch reduction order is not specified for > 2 ranks
        out.append(torch.cat([v["master"].reshape(-1) for _, v in sorted(snap_opt["state"].items())]))
    return torch.cat(out)


def test_fused_sharded_optimizer_state_dict_roundtrip():
    res = run_distributed(_sharded_state_worker, world=_ngpu(), use_cuda=True)
    for r in res[1:]:
        assert torch.equal(res[0], r)


# ---------------------------------------------------------------------------------------------------------------------
# Decentralized algorithms against python re-implementations on real GPUs, with a bucket large enough (>= 8 MiB) to leave the
# one-shot / padding paths (reference: tests/torch_api/test_decentralized.py:169-259,326-399 and
# tests/torch_api/test_low_precision_decentralized.py:157-255).
# ---------------------------------------------------------------------------------------------------------------------
def _big_mlp(dev):
    torch.manual_seed(77)
    return torch.nn.Sequential(torch.nn.Linear(1024, 1024), torch.nn.ReLU(), torch.nn.Linear(1024, 1024), torch.nn.ReLU(), torch.nn.Linear(1024, 64)).to(dev)


def _exchange(t, peer):
    import torch.distributed as dist

    got = torch.empty_like(t)
    for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, t, peer), dist.P2POp(dist.irecv, got, peer)]):
        r.wait()
    return got


def _ring_exchange(t, left, right):
    """Send ``t`` to both ring neighbours and receive theirs, as ONE batch (a send-to-left paired with a receive-from-left only
    matches up on every rank when the world is 2)."""
    import torch.distributed as dist

    from_left, from_right = torch.empty_like(t), torch.empty_like(t)
    if left == right:
        ops = [dist.P2POp(dist.isend, t, left), dist.P2POp(dist.irecv, from_left, left)]
    else:
        ops = [dist.P2POp(dist.isend, t, left), dist.P2POp(dist.isend, t, right), dist.P2POp(dist.irecv, from_left, left), dist.P2POp(dist.irecv, from_right, right)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    return from_left, (from_left if left == right else from_right)


def _decentralized_oracle_worker(rank, world, mode):
    import copy

    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.core import native
    from bagua_b200.parallel.algorithms import decentralized

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    model = _big_mlp(dev)
    oracle = copy.deepcopy(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    oopt = torch.optim.SGD(oracle.parameters(), lr=0.05)
    model = model.with_bagua([opt], decentralized.DecentralizedAlgorithm(hierarchical=False, peer_selection_mode=mode))
    bucket = model.bagua_buckets[0]
    assert bucket.bytes() >= 8 * 1024 * 1024
    assert "python" not in bucket.backend_bucket.print_ops(), bucket.backend_bucket.print_ops()
    oparams = list(reversed(list(oracle.parameters())))      # bucket order
    for step in range(4):
        x = torch.randn(16, 1024, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * step + rank))
        # --- oracle: average the weights as they are at forward time, swap them in before the optimizer step
        flat = torch.cat([p.data.reshape(-1) for p in oparams])
        if mode == "all":
            pw = flat.clone()
            dist.all_reduce(pw)
            pw /= world
        else:
            peer = native().PeerAverageOp.shift_one_peer(rank, world, step)
            pw = (flat + _exchange(flat, peer)) / 2
        oopt.zero_grad()
        oracle(x).pow(2).mean().backward()
        off = 0
        with torch.no_grad():
            for p in oparams:
                p.copy_(pw[off: off + p.numel()].view_as(p))
                off += p.numel()
        oopt.step()
        # --- bagua
        opt.zero_grad()
        model(x).pow(2).mean().backward()
        torch.cuda.synchronize()
        got_pw = bucket._peer_weight.view(-1)[: pw.numel()]
        torch.testing.assert_close(got_pw, pw, rtol=1e-6, atol=1e-6)           # the peer_weight replica itself
        opt.step()
    torch.cuda.synchronize()
    mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    want = torch.cat([p.detach().reshape(-1) for p in oracle.parameters()])
    torch.testing.assert_close(mine, want, rtol=1e-5, atol=1e-6)
    assert bagua.communication._get_default_group().peer_engine().comm.error_code() == 0
    return True


@pytest.mark.parametrize("mode", ["all", "shift_one"])
def test_decentralized_matches_python_oracle_on_gpus(mode):
    world = _ngpu() if _ngpu() % 2 == 0 else _ngpu() - 1
    run_distributed(_decentralized_oracle_worker, world=world, args=(mode,), use_cuda=True)


def _lp_decentralized_oracle_worker(rank, world):
    import copy

    import bagua_b200 as bagua
    from bagua_b200.ops import quant
    from bagua_b200.parallel.algorithms import decentralized

    bagua.init_process_group()
    dev = torch.device("cuda", rank)
    model = _big_mlp(dev)
    oracle = copy.deepcopy(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    oopt = torch.optim.SGD(oracle.parameters(), lr=0.05)
    model = model.with_bagua([opt], decentralized.LowPrecisionDecentralizedAlgorithm(hierarchical=False))
    assert all("low_precision_ring_fused" in b.backend_bucket.print_ops() for b in model.bagua_buckets)
    assert sum(b.bytes() for b in model.bagua_buckets) >= 8 * 1024 * 1024
    # the oracle keeps weight / left / right replicas per bucket, in bucket layout
    names = {id(p): n for n, p in model.named_parameters()}
    oparam = dict(oracle.named_parameters())
    state = []
    for b in model.bagua_buckets:
        ps = [oparam[names[id(t)]] for t in b.tensors]
        w = torch.cat([p.data.reshape(-1) for p in ps])
        state.append({"params": ps, "w": w.clone(), "l": w.clone(), "r": w.clone()})
    left, right = (rank + world - 1) % world, (rank + 1) % world
    for step in range(4):
        x = torch.randn(16, 1024, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * step + rank))
        oopt.zero_grad()
        oracle(x).pow(2).mean().backward()
        oopt.step()
        for st in state:
            xx = torch.cat([p.data.reshape(-1) for p in st["params"]])
            xx.add_(st["l"], alpha=1.0 / 3.0).add_(st["r"], alpha=1.0 / 3.0).sub_(st["w"], alpha=5.0 / 3.0)
            mm, q = quant.torch_compress_chunk(xx)
            (lmm, rmm), (lq, rq) = _ring_exchange(mm, left, right), _ring_exchange(q, left, right)
            st["l"] += quant.torch_decompress_chunk(lmm, lq, xx.dtype)
            st["r"] += quant.torch_decompress_chunk(rmm, rq, xx.dtype)
            xx = st["w"] + quant.torch_decompress_chunk(mm, q, xx.dtype)
            st["w"].copy_(xx)
            off = 0
            with torch.no_grad():
                for p in st["params"]:
                    p.copy_(xx[off: off + p.numel()].view_as(p))
                    off += p.numel()
        opt.zero_grad()
        model(x).pow(2).mean().backward()
        opt.step()
    torch.cuda.synchronize()
    for b, st in zip(model.bagua_buckets, state):
        n = st["w"].numel()
        for got, want, what in ((b._weight.view(-1)[:n], st["w"], "weight"), (b._left_peer_weight.view(-1)[:n], st["l"], "left"),
                                (b._right_peer_weight.view(-1)[:n], st["r"], "right")):
            level = (want.max() - want.min()).item() / 255
            d = (got - want).abs()
            assert d.max().item() <= 4 * level + 1e-5, (what, d.max().item(), level)       # rare one-level flips compound over 4 steps
            assert (d > 1e-4).float().mean().item() < 1e-2, what
    assert bagua.communication._get_default_group().peer_engine().comm.error_code() == 0
    return True


def test_low_precision_decentralized_replicas_match_python_oracle_on_gpus():
    run_distributed(_lp_decentralized_oracle_worker, world=_ngpu(), use_cuda=True)


def _fused_sgd_oracle_worker(rank, world, variant):
    """The kernel behind the N > 1 headline: reduce-scatter → SGD(momentum, nesterov, wd) → all-gather vs torch.optim.SGD on
    all-reduced gradients, bf16 and fp32, two-shot (peer ld/st) and multimem (NVLS) flavours, buckets from 64 KiB to 256 MiB."""
    import torch.distributed as dist

    import bagua_b200 as bagua
    from bagua_b200.core import dtype_code, native

    bagua.init_process_group()
    eng = bagua.communication._get_default_group().peer_engine()
    C = native()
    dev = torch.device("cuda", rank)
    use_mc = variant == "multimem"
    if use_mc and not eng.has_multicast:
        return "no multicast"
    stream = torch.cuda.current_stream().cuda_stream
    for dtype in (torch.float32, torch.bfloat16):
        es = torch.empty(0, dtype=dtype).element_size()
        for nbytes in (64 * 1024, 8 * 1024 ** 2 + 48, 256 * 1024 ** 2):
            nbytes = nbytes // 16 * 16
            numel = nbytes // es
            gs, ws = eng.alloc(nbytes), eng.alloc(nbytes)
            per = 16 // es
            vpr = (nbytes // 16 + world - 1) // world
            lo, hi = rank * vpr * per, min((rank + 1) * vpr * per, numel)
            torch.manual_seed(1)
            w0 = torch.randn(numel, device=dev).to(dtype)
            ws.view(dtype, numel).copy_(w0)
            master = torch.zeros(vpr * per, device=dev)
            if hi > lo:
                master[: hi - lo].copy_(w0[lo:hi].float())
            mom = torch.zeros(vpr * per, device=dev)
            op = C.AllReduceSgdOp(eng.comm, gs.buf, ws.buf, gs.offset, ws.offset, nbytes, dtype_code(dtype), master.data_ptr(), mom.data_ptr(), 1.0 / world, True,
                                  use_mc, eng.launch_cfg("multimem" if use_mc else "two_shot", nbytes, 16 if use_mc else 32))
            op.set_hyper(0.1, 0.9, 0.0, 1e-4, True)
            ref = torch.nn.Parameter(w0.float().clone())
            ropt = torch.optim.SGD([ref], lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
            for step in range(3):
                torch.manual_seed(10 * step + rank)
                g = torch.randn(numel, device=dev).to(dtype)
                gs.view(dtype, numel).copy_(g)
                gsum = g.float()
                dist.all_reduce(gsum)
                ref.grad = gsum / world
                ropt.step()
                torch.cuda.synchronize()
                dist.barrier()
                C.run_op(op, stream, rank)
                torch.cuda.synchronize()
                assert gs.view(dtype, numel).abs().max().item() == 0.0
            if hi > lo:
                tol = 1e-4 if dtype == torch.float32 else 2e-2      # bf16: the in-switch sum is rounded to bf16 once
                torch.testing.assert_close(master[: hi - lo], ref.data[lo:hi], rtol=tol, atol=tol)
            got = ws.view(dtype, numel).float()
            torch.testing.assert_close(got, ref.data.to(dtype).float(), rtol=0, atol=4e-2 if dtype == torch.bfloat16 else 1e-4)
            gs.free()
            ws.free()
    assert eng.comm.error_code() == 0
    return "ok"


@pytest.mark.parametrize("variant", ["two_shot", "multimem"])
def test_fused_allreduce_sgd_kernel_matches_torch_sgd(variant):
    res = run_distributed(_fused_sgd_oracle_worker, world=_ngpu(), args=(variant,), use_cuda=True, timeout=600)
    assert all(r in ("ok", "no multicast") for r in res)
