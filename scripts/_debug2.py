import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bagua_b200.core import native, dtype_code
from bagua_b200.parallel.virtual import VirtualPeerWorld
print("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
C = native()
for variant in ("release_only", "add_then_release", "sleepkernel"):
    w = VirtualPeerWorld(1, dev, timeout_s=20.0)
    numel = 1 << 14; nbytes = numel * 4
    snap, avg = w.alloc(nbytes), w.alloc(nbytes)
    wt = torch.zeros(numel, device=dev)
    gate = C.WeightGate(0)
    op = C.AsyncAverageOp(w.comms[0], wt.data_ptr(), snap.buf, 0, avg.buf, 0, nbytes, dtype_code(torch.float32), gate, 4.0, False, w.cfg(2))
    trainer = torch.cuda.Stream()
    other = torch.zeros(1024, device=dev)
    gate.acquire(trainer.cuda_stream, 1.0); torch.cuda.synchronize()
    t0 = time.time()
    C.run_op(op, w.streams[0].cuda_stream, 0)
    time.sleep(0.1)
    ev = torch.cuda.Event()
    with torch.cuda.stream(trainer):
        if variant == "add_then_release":
            other.add_(1.0)
        if variant == "sleepkernel":
            torch.cuda._sleep(1000000)
        ev0 = torch.cuda.Event(); ev0.record()
        gate.release(trainer.cuda_stream)
        ev.record()
    t_ev0 = t_ev = t_k = None
    while time.time() - t0 < 8:
        now = time.time() - t0
        if t_ev0 is None and ev0.query(): t_ev0 = now
        if t_ev is None and ev.query(): t_ev = now
        if t_k is None and w.streams[0].query(): t_k = now
        if t_ev is not None and t_k is not None: break
        time.sleep(0.001)
    torch.cuda.synchronize()
    print(variant, "pre-release event at", t_ev0, "release done at", t_ev, "avg kernel done at", t_k, "status", op.status(), "w", wt[0].item(), "gate", gate.state())
