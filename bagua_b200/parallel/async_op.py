"""The asynchronous model-average bucket op (comm op 5 of SURVEY §2.5; reference
rust/bagua-core/bagua-core-internal/src/comm_ops/decentralized_full_precision_asynchronous.rs:30-180).

One averaging round, executed on the scheduler's worker thread:

1. all ranks agree whether anybody asked to stop (MIN-allreduce of a flag on the background group);
2. snapshot the weights; 3. sum the snapshots over all ranks — on NVSwitch one out-of-place two-shot / multimem kernel
   from the symmetric ``snap`` buffer into the symmetric ``red`` buffer, on the background group's own stream;
4. under the weight lock apply ``w += red/P − snap`` (one kernel, on the compute stream so it is ordered with the
   optimizer step exactly like the reference).
"""
from __future__ import annotations

import threading

import torch
import torch.distributed as dist

from ..core import dtype_code, native


class AsyncModelAverageOp:
    def __init__(self, bucket, group):
        self.bucket = bucket
        self.group = group
        self.lock = threading.Lock()
        self._abort = False
        self._running = True
        self._trainer_holds = False
        flat = bucket.backend_tensor
        assert flat is not None, "Async algorithm supports `do_flatten=True` only"
        self.flat = flat
        self.snap = bucket.new_companion(init_from_bucket=False, symmetric=True, group=group)
        self.red = bucket.new_companion(init_from_bucket=False, symmetric=True, group=group)
        self._native = None
        eng = bucket._engine(group)
        s_snap, s_red = getattr(self.snap, "_bagua_symm_slice", None), getattr(self.red, "_bagua_symm_slice", None)
        if eng is not None and s_snap is not None and s_red is not None and flat.dtype in (torch.float32, torch.float16, torch.bfloat16):
            nbytes = flat.numel() * flat.element_size()
            if nbytes % 16 == 0:
                op, _ = eng.make_allreduce_op(s_snap, s_red, nbytes, flat.dtype, False, "auto")
                self._native = (eng, op)
        dev = flat.device
        self._flag = torch.ones(1, dtype=torch.int32, device=dev)

    # -- mutex protocol (trainer holds it from forward-pre to post-backward) ---------------------------------------
    # Only the training thread calls lock_weight / unlock_weight; the comm worker takes ``self.lock`` around the apply.
    # A forward that is never followed by a backward (evaluation in train mode) leaves the trainer holding the lock;
    # the next forward must not dead-lock on it, hence the ownership flag instead of a bare acquire.
    def lock_weight(self):
        if self._trainer_holds:
            return
        self.lock.acquire()
        self._trainer_holds = True

    def unlock_weight(self):
        if self._trainer_holds:
            self._trainer_holds = False
            self.lock.release()

    def abort(self):
        self._abort = True

    def reset(self):
        self._abort = False
        self._running = True

    def get_status(self) -> bool:
        return self._running

    # -- one round ---------------------------------------------------------------------------------------------------
    def run(self, _bucket_name: str):
        pg = self.group
        n = pg.size()
        cuda = self.flat.is_cuda
        with torch.no_grad():
            self._flag.fill_(0 if self._abort else 1)
            if n > 1:
                if cuda:
                    with torch.cuda.stream(pg.stream):
                        dist.all_reduce(self._flag, op=dist.ReduceOp.MIN, group=pg.torch_group)
                else:
                    dist.all_reduce(self._flag, op=dist.ReduceOp.MIN, group=pg.torch_group)
            if int(self._flag.item()) == 0:
                self._running = False
                return
            if cuda:
                main = torch.cuda.current_stream()
                self.snap.copy_(self.flat)  # on the compute stream, like the reference (…asynchronous.rs:124)
                ev = main.record_event()
                pg.stream.wait_event(ev)
                if self._native is not None:
                    eng, op = self._native
                    native().run_op(op, pg.stream.cuda_stream, self.flat.device.index)
                else:
                    with torch.cuda.stream(pg.stream):
                        self.red.copy_(self.snap)
                        if n > 1:
                            dist.all_reduce(self.red, group=pg.torch_group)
                pg.stream.synchronize()
                with self.lock:
                    native().async_apply(self.flat.data_ptr(), self.red.data_ptr(), self.snap.data_ptr(), self.flat.numel(), dtype_code(self.flat.dtype),
                                         1.0 / n, main.cuda_stream)
                    main.synchronize()
            else:
                self.snap.copy_(self.flat)
                self.red.copy_(self.snap)
                if n > 1:
                    dist.all_reduce(self.red, group=pg.torch_group)
                with self.lock:
                    self.flat.add_(self.red / n - self.snap)


class FusedAsyncModelAverageOp:
    """GPU flavour of the op: one ``async_average_kernel`` launch per round (vote + snapshot → mean of the snapshots over peer
    memory → ``w += mean − snapshot``), issued by the scheduler's worker as a native op — no python on the worker, no NCCL,
    no ``.item()``, no stream synchronisation.

    The reference's host mutex (held by the trainer from forward-pre to post-backward, with a host sync at both ends,
    algorithms/async_model_average.py:212-225) becomes a *device-side weight gate* (``csrc``: ``WeightGate``): the trainer's
    stream acquires it before a forward pass and releases it after the optimizer step; the averaging kernel takes it only for
    its final, purely local apply phase.  ``lock_weight`` / ``unlock_weight`` keep their names; they enqueue the gate kernels on
    the caller's current stream and return immediately."""

    def __init__(self, bucket, group, eng, snap_slice, avg_slice):
        C = native()
        flat = bucket.backend_tensor
        self.bucket, self.group, self.flat = bucket, group, flat
        self.snap, self.red = snap_slice, avg_slice
        nbytes = flat.numel() * flat.element_size()
        self.gate = C.WeightGate(flat.device.index)
        use_mc = bool(eng.has_multicast and snap_slice.has_multicast and avg_slice.has_multicast and eng.world > 2)
        import os

        self.gate_timeout_s = float(os.environ.get("BAGUA_ASYNC_GATE_TIMEOUT_S", "2.0"))
        blocks = int(os.environ.get("BAGUA_ASYNC_BLOCKS", "0")) or 16
        self.native_op = C.AsyncAverageOp(eng.comm, flat.data_ptr(), snap_slice.buf, snap_slice.offset, avg_slice.buf, avg_slice.offset, nbytes,
                                          dtype_code(flat.dtype), self.gate, self.gate_timeout_s, use_mc, eng.launch_cfg("two_shot", nbytes * 4, blocks))
        self.variant = "async_fused_" + ("multimem" if use_mc else "peer")
        self._trainer_holds = False
        self._stopped = False

    @classmethod
    def create(cls, bucket, group):
        flat = bucket.backend_tensor
        assert flat is not None, "Async algorithm supports `do_flatten=True` only"
        eng = bucket._engine(group)
        if eng is None or not flat.is_cuda or flat.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            return None
        nbytes = flat.numel() * flat.element_size()
        if nbytes % 16:
            return None
        snap, avg = eng.alloc(nbytes), eng.alloc(nbytes)
        bucket._companion_slices += [snap, avg]
        return cls(bucket, group, eng, snap, avg)

    # trainer side (training thread, its current stream) ---------------------------------------------------------------
    def lock_weight(self):
        if not self._trainer_holds:
            self.gate.acquire(torch.cuda.current_stream().cuda_stream, max(self.gate_timeout_s * 5, 10.0))
            self._trainer_holds = True

    def unlock_weight(self):
        if self._trainer_holds:
            self.gate.release(torch.cuda.current_stream().cuda_stream)
            self._trainer_holds = False

    # control ------------------------------------------------------------------------------------------------------------
    def abort(self):
        self.native_op.abort()

    def reset(self):
        self.native_op.reset()
        self._stopped = False

    def get_status(self) -> bool:
        """False once a completed round reported that some rank voted to stop (every rank sees the same round outcome)."""
        st = self.native_op.status()
        if st < 0:
            raise RuntimeError("bagua: asynchronous model averaging failed (peer barrier time-out or protocol violation)")
        if st == 0:
            self._stopped = True
        return not self._stopped
