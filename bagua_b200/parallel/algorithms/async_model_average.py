"""Asynchronous model averaging (reference: bagua/torch_api/algorithms/async_model_average.py:1-347).

Workers train without waiting for each other; a background loop averages the model every ``sync_interval_ms`` on a
dedicated process group (own stream, own signal pads) while the trainer holds a weight lock only from forward-pre to
post-backward.  Optional gradient-allreduce warm-up.  ``abort()``/``resume()`` stop/restart the loop on all ranks."""
from __future__ import annotations

import logging
import threading
import time
from typing import Callable, List, Optional

import torch

from ... import communication as comm_mod
from ...bucket import BaguaBucket
from .base import Algorithm, AlgorithmImpl

__all__ = ["AsyncModelAverageAlgorithm", "AsyncModelAverageAlgorithmImpl"]

logger = logging.getLogger(__name__)


class _AveragingLoop:
    """Host-side driver of the averaging rounds.  Life cycle: ``idle`` → :meth:`arm` (thread exists, parked) → :meth:`release` (first
    training forward: rounds begin) → the ``keep_going`` predicate turns false (the op was aborted) → :meth:`join` → ``stopped``;
    ``arm`` again restarts it.  One short-lived daemon thread per arm, parked on an event — nothing polls while training is idle."""

    def __init__(self, one_round: Callable[[int], None], keep_going: Callable[[], bool], interval_s: float, on_thread_start: Callable[[], None]):
        self._one_round, self._keep_going, self._interval_s, self._on_thread_start = one_round, keep_going, interval_s, on_thread_start
        self._go = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.state = "idle"

    def arm(self):
        self._go.clear()
        self._thread = threading.Thread(target=self._main, name="bagua-async-average", daemon=True)
        self._thread.start()
        self.state = "armed"

    def release(self):
        self._go.set()
        if self.state == "armed":
            self.state = "running"

    def join(self):
        self._go.set()                      # an armed loop that never saw a forward must still be able to leave
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        self.state = "stopped"

    @property
    def live(self) -> bool:
        return self.state in ("armed", "running")

    def _main(self):
        self._go.wait()
        self._on_thread_start()
        rounds = 0
        while self._keep_going():
            t0 = time.time()
            self._one_round(rounds)
            logger.debug("async communication cost %.2f ms, comm_step=%d", (time.time() - t0) * 1000, rounds)
            rounds += 1
            time.sleep(self._interval_s)


def _unwrap_ddp(obj):
    from ..bagua_distributed import BaguaDistributedDataParallel

    if isinstance(obj, BaguaDistributedDataParallel):
        return obj
    if hasattr(obj, "inner"):
        return obj.inner
    if hasattr(obj, "bagua_ddp"):
        return obj.bagua_ddp
    raise Exception(f"Unexpect input bagua_ddp({type(obj)}), it should be BaguaDistributedDataParallel or a module wrapped by with_bagua.")


class AsyncModelAverageAlgorithmImpl(AlgorithmImpl):
    def __init__(self, process_group, peer_selection_mode: str = "all", sync_interval_ms: int = 500, warmup_steps: int = 0):
        super().__init__(process_group)
        self.peer_selection_mode, self.sync_interval_ms, self.warmup_steps = peer_selection_mode, sync_interval_ms, warmup_steps
        self.step_id = 0
        self._loop: Optional[_AveragingLoop] = None      # created with the first engine that needs it (see _loop_for)
        # background communication gets its own group: own comm stream, own NCCL/gloo communicator, own signal pads
        self.thread_group = comm_mod.new_group(list(process_group.ranks), stream=comm_mod._make_stream())

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        assert do_flatten, "Async algorithm supports `do_flatten=True` only"
        if self.step_id < self.warmup_steps:
            return [BaguaBucket(b, flatten=do_flatten, name=str(i), group=self.thread_group) for i, b in enumerate(tensors)]
        all_tensors = [t for b in tensors for t in b]
        align = max(1, 16 // all_tensors[0].element_size()) * self.process_group.size()
        return [BaguaBucket(all_tensors, flatten=do_flatten, name="0", alignment=align, group=self.thread_group)]

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        """Warm-up steps communicate gradients (plain all-reduce); afterwards the weights themselves are the bucket."""
        named = bagua_ddp.bagua_build_params()
        self._communication_tensor_names = {n for n, _ in named}
        owner = bagua_ddp.bagua_module_name
        if self.step_id >= self.warmup_steps:
            return [p.ensure_bagua_tensor(n, owner) for n, p in reversed(named)]

        def grad_of(p):
            return p.grad

        def set_grad(p, t):
            p.grad = t

        return [p.bagua_ensure_grad().ensure_bagua_tensor(n, owner, getter_closure=grad_of, setter_closure=set_grad) for n, p in reversed(named)]

    def _averaging_phase(self) -> bool:
        return self.step_id > self.warmup_steps and self.sync_interval_ms > 0

    def _loop_for(self, bagua_ddp) -> _AveragingLoop:
        if self._loop is None:
            def one_round(_n):
                for bucket in bagua_ddp.bagua_buckets:
                    for t in bucket.tensors:
                        t.bagua_mark_communication_ready_without_synchronization()
                bagua_ddp._bagua_backend.wait_pending_comm_ops(0, True)

            def bind_device():
                if torch.cuda.is_available() and comm_mod._use_cuda():
                    torch.cuda.set_device(comm_mod._device_index())

            self._loop = _AveragingLoop(one_round, lambda: self._check_op_status(bagua_ddp), self.sync_interval_ms / 1000, bind_device)
        return self._loop

    def init_forward_pre_hook(self, bagua_ddp):
        def take_weights(_inputs):
            if not self._averaging_phase():
                return
            loop = self._loop_for(bagua_ddp)
            if loop.state == "idle":         # the user never called resume(): the first training forward starts the rounds
                loop.arm()
            loop.release()
            self._lock_model(bagua_ddp)      # the trainer owns the weights from here to post-backward (or to the optimizer step)

        return take_weights

    def init_backward_hook(self, bagua_ddp):
        def warmup_gradient_ready(parameter_name, parameter):
            if self.step_id <= self.warmup_steps and parameter_name in self._communication_tensor_names:
                bagua_ddp.mark_tensor_ready(parameter)

        return warmup_gradient_ready

    def init_post_backward_hook(self, bagua_ddp):
        def after_backward():
            if self.step_id <= self.warmup_steps:
                bagua_ddp.wait_pending_comm_ops()
            elif not getattr(self, "_release_after_step", False):
                self._unlock_model(bagua_ddp)

        return after_backward

    def need_reset(self) -> bool:
        self.step_id += 1
        switching = self.warmup_steps > 0 and self.step_id == self.warmup_steps + 1    # gradient buckets → one weight bucket
        if switching:
            logger.info("Async model average starts from step %d", self.step_id)
        return switching

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        bucket.clear_ops()
        if self.step_id < self.warmup_steps:
            bucket.append_centralized_synchronous_op(hierarchical=False, average=True, group=self.thread_group)
        else:
            bucket._async_op = bucket.append_asynchronous_model_average_op(peer_selection_mode=self.peer_selection_mode, group=self.thread_group)
            self._install_step_hooks(bagua_ddp)

    def _sync_compute_stream(self):
        if torch.cuda.is_available() and comm_mod._use_cuda():
            torch.cuda.current_stream().synchronize()

    @staticmethod
    def _gated(bagua_ddp) -> bool:
        """True when every bucket runs the fused kernel: the weight lock is then a device-side gate (stream-ordered, no host
        sync); otherwise it is the host mutex of the reference and the compute stream is drained around it."""
        from ..async_op import FusedAsyncModelAverageOp

        ops = [getattr(b, "_async_op", None) for b in bagua_ddp.bagua_buckets]
        return bool(ops) and all(isinstance(o, FusedAsyncModelAverageOp) for o in ops)

    @staticmethod
    def _async_ops(bagua_ddp):
        return [b._async_op for b in bagua_ddp.bagua_buckets if hasattr(b, "_async_op")]

    def _lock_model(self, bagua_ddp):
        if not self._gated(bagua_ddp):
            self._sync_compute_stream()
        for op in self._async_ops(bagua_ddp):
            op.lock_weight()

    def _unlock_model(self, bagua_ddp):
        if not self._gated(bagua_ddp):
            self._sync_compute_stream()
        for op in self._async_ops(bagua_ddp):
            op.unlock_weight()

    def _install_step_hooks(self, bagua_ddp):
        """With the device-side gate the trainer keeps the weights until the optimizer has stepped (so the averaging delta never
        interleaves with the optimizer's read-modify-write): the release moves from post-backward into a step post-hook of the
        optimizers handed to ``with_bagua``.  Without optimizers (or on the host-mutex path) the reference's post-backward unlock stays."""
        self._release_after_step = False
        if not self._gated(bagua_ddp):
            return
        opts = [o for o in getattr(bagua_ddp, "bagua_optimizers", []) if hasattr(o, "register_step_post_hook")]
        if not opts:
            return
        for o in opts:
            if getattr(o, "_bagua_async_gate_hook", None) is not None:
                o._bagua_async_gate_hook.remove()
            o._bagua_async_gate_hook = o.register_step_post_hook(lambda *_a, **_k: self._unlock_model(bagua_ddp))
        self._release_after_step = True

    def _check_op_status(self, bagua_ddp) -> bool:
        ops = self._async_ops(bagua_ddp)
        return bool(ops) and bool(ops[0].get_status())

    def abort(self, bagua_ddp):
        """Stop the background averaging on every rank (call after training / before evaluation)."""
        bagua_ddp = _unwrap_ddp(bagua_ddp)
        loop = self._loop_for(bagua_ddp)
        if not loop.live:
            return
        self._unlock_model(bagua_ddp)  # a forward without backward must not keep the averaging round (and this abort) waiting
        comm_mod.barrier(comm=self.process_group.get_global_communicator())
        for op in self._async_ops(bagua_ddp)[:1]:
            op.abort()                  # the rounds' own vote makes every rank leave after the same round
        loop.join()
        logger.debug("async communication aborted.")

    def resume(self, bagua_ddp):
        """Restart the background averaging stopped by :meth:`abort` (call before training)."""
        bagua_ddp = _unwrap_ddp(bagua_ddp)
        loop = self._loop_for(bagua_ddp)
        if loop.live:
            return
        comm_mod.barrier(comm=self.process_group.get_global_communicator())
        for op in self._async_ops(bagua_ddp)[:1]:
            op.reset()
        loop.arm()
        logger.debug("async communication resumed.")


class AsyncModelAverageAlgorithm(Algorithm):
    def __init__(self, peer_selection_mode: str = "all", sync_interval_ms: int = 500, warmup_steps: int = 0):
        """
        Args:
            peer_selection_mode: only ``"all"`` (every worker's weights are averaged each round).
            sync_interval_ms: milliseconds between two averaging rounds.
            warmup_steps: gradient-allreduce steps before asynchronous averaging starts (0 disables the warm-up).
        """
        self.peer_selection_mode, self.sync_interval_ms, self.warmup_steps = peer_selection_mode, sync_interval_ms, warmup_steps

    def reify(self, process_group) -> AsyncModelAverageAlgorithmImpl:
        return AsyncModelAverageAlgorithmImpl(process_group, self.peer_selection_mode, self.sync_interval_ms, self.warmup_steps)
