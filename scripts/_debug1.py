import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BAGUA_SELF_PEER"] = "1"
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29877")
import torch
import bagua_b200 as bagua
from bagua_b200.core import native, dtype_code
torch.cuda.set_device(0)
bagua.init_process_group()
dev = torch.device("cuda", 0)
C = native()
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "fused"):
    from bagua_b200.parallel.algorithms.gradient_allreduce import FusedGradientAllReduceAlgorithm, make_sharded_fused_sgd
    torch.manual_seed(7)
    model = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 64)).to(dev)
    oracle = copy.deepcopy(model)
    opt = make_sharded_fused_sgd(model.parameters(), lr=0.05)
    oopt = torch.optim.SGD(oracle.parameters(), lr=0.05)
    model = model.with_bagua([opt], FusedGradientAllReduceAlgorithm(opt))
    def diff():
        return max((p.detach() - q.detach()).abs().max().item() for p, q in zip(model.parameters(), oracle.parameters()))
    print("after with_bagua: diff", diff(), "buckets", [(b.name, b.allreduce_variant, b.backend_bucket.print_ops()) for b in model.bagua_buckets])
    for rec in opt._shards:
        print("shard", rec["bucket"], rec["lo"], rec["hi"], rec["numel"], "master absmax", rec["state"][0].abs().max().item(), "weights absmax", rec["weights"].abs().max().item())
    for it in range(3):
        x = torch.randn(32, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(it))
        opt.zero_grad(); model(x).pow(2).mean().backward()
        torch.cuda.synchronize()
        model.bagua_ddp.wait_pending_comm_ops(); torch.cuda.synchronize()
        print("step", it, "grad absmax after bwd (should be 0, cleared by kernel)", max(p.grad.abs().max().item() for p in model.parameters()),
              "weights absmax", max(p.detach().abs().max().item() for p in model.parameters()), "master absmax", [r["state"][0].abs().max().item() for r in opt._shards],
              "op steps", [r["op"].steps() for r in opt._shards], "err", bagua.communication._get_default_group().peer_engine().comm.error_code())
        opt.step()
        oopt.zero_grad(); oracle(x).pow(2).mean().backward(); oopt.step()
        print("   diff vs oracle", diff())

if which in ("all", "gate"):
    from bagua_b200.parallel.virtual import VirtualPeerWorld
    P = 2
    w = VirtualPeerWorld(P, dev, timeout_s=20.0)
    numel = 1 << 14; nbytes = numel * 4
    snap, avg = w.alloc(nbytes), w.alloc(nbytes)
    weights = [torch.full((numel,), float(r), device=dev) for r in range(P)]
    gates = [C.WeightGate(0) for _ in range(P)]
    ops = [C.AsyncAverageOp(w.comms[r], weights[r].data_ptr(), snap.buf, 0, avg.buf, 0, nbytes, dtype_code(torch.float32), gates[r], 10.0, False, w.cfg(2)) for r in range(P)]
    trainer = torch.cuda.Stream()
    gates[0].acquire(trainer.cuda_stream, 1.0); torch.cuda.synchronize()
    print("gate0 state", gates[0].state())
    t0 = time.time()
    for r in range(P):
        C.run_op(ops[r], w.streams[r].cuda_stream, 0)
    time.sleep(0.2)
    print("rank1 stream done?", w.streams[1].query(), "rank0 stream done?", w.streams[0].query(), "w1[0]", "n/a")
    with torch.cuda.stream(trainer):
        weights[0].add_(100.0)
        gates[0].release(trainer.cuda_stream)
    torch.cuda.synchronize()
    print("elapsed", time.time() - t0, "w0", weights[0][:3].tolist(), weights[0][-3:].tolist(), "w1", weights[1][:3].tolist(), "status", [o.status() for o in ops],
          "snap0", snap.view(0, torch.float32)[:2].tolist(), "avg0", avg.view(0, torch.float32)[:2].tolist(), avg.view(0, torch.float32)[-2:].tolist(), "gate", gates[0].state(),
          "err", [c.error_code() for c in w.comms])

if which in ("all", "qadam"):
    from bagua_b200.parallel.algorithms import q_adam
    import torch.nn as nn
    torch.manual_seed(3)
    base = nn.Sequential(nn.Linear(64, 256), nn.ReLU(), nn.Linear(256, 32)).to(dev)
    model, oracle = copy.deepcopy(base), copy.deepcopy(base)
    opt = q_adam.QAdamOptimizer(model.parameters(), lr=1e-3, warmup_steps=3)
    oopt = q_adam.QAdamOptimizer(oracle.parameters(), lr=1e-3, warmup_steps=3)
    model = model.with_bagua([opt], q_adam.QAdamAlgorithm(opt))
    for it in range(8):
        x = torch.randn(16, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(it))
        opt.zero_grad(); model(x).square().mean().backward(); opt.step()
        oopt.zero_grad(); oracle(x).square().mean().backward()
        if it >= 3:
            for p in oracle.parameters():
                oopt.state[p]["exp_avg"].mul_(0.9).add_(p.grad, alpha=0.1)
        oopt.step()
        torch.cuda.synchronize()
        md = max((opt.state[p]["exp_avg"] - oopt.state[q]["exp_avg"]).abs().max().item() for p, q in zip(model.parameters(), oracle.parameters()))
        wd = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(model.parameters(), oracle.parameters()))
        gd = max((p.grad - q.grad).abs().max().item() for p, q in zip(model.parameters(), oracle.parameters()))
        print("it", it, "ops", [b.backend_bucket.print_ops() for b in model.bagua_buckets], "exp_avg diff", md, "weight diff", wd, "grad diff", gd,
              "mom absmax", max(opt.state[p]["exp_avg"].abs().max().item() for p in model.parameters()))
