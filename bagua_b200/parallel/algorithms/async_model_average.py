"""Asynchronous model averaging (reference: bagua/torch_api/algorithms/async_model_average.py:1-347).

Workers train without waiting for each other; a background loop averages the model every ``sync_interval_ms`` on a
dedicated process group (own stream, own signal pads) while the trainer holds a weight lock only from forward-pre to
post-backward.  Optional gradient-allreduce warm-up.  ``abort()``/``resume()`` stop/restart the loop on all ranks."""
from __future__ import annotations

import logging
import threading
import time
from concurrent.futures import ThreadPoolExecutor, wait
from enum import IntEnum
from typing import List

import torch

from ... import communication as comm_mod
from ...bucket import BaguaBucket
from .base import Algorithm, AlgorithmImpl

__all__ = ["AsyncModelAverageAlgorithm", "AsyncModelAverageAlgorithmImpl"]

logger = logging.getLogger(__name__)


class _AsyncInternalState(IntEnum):
    NEW = 0
    SCHEDULED = 1
    STARTED = 2
    STOPPED = 3


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
        self.peer_selection_mode = peer_selection_mode
        self.sync_interval_ms = sync_interval_ms
        self.step_id = 0
        self.warmup_steps = warmup_steps
        self.executor = ThreadPoolExecutor(max_workers=1)
        self.cv = threading.Condition()
        self.notified = False
        self.status = _AsyncInternalState.NEW
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
        parameters = bagua_ddp.bagua_build_params()
        tensors = []
        for name, param in reversed(parameters):
            if self.step_id < self.warmup_steps:
                t = param.bagua_ensure_grad().ensure_bagua_tensor(
                    name, bagua_ddp.bagua_module_name, getter_closure=lambda p: p.grad, setter_closure=lambda p, t: setattr(p, "grad", t)
                )
            else:
                t = param.ensure_bagua_tensor(name, bagua_ddp.bagua_module_name)
            tensors.append(t)
        self._communication_tensor_names = set(name for name, _ in parameters)
        return tensors

    def init_forward_pre_hook(self, bagua_ddp):
        def hook(input):
            if self.step_id > self.warmup_steps and self.sync_interval_ms > 0:
                if self.status == _AsyncInternalState.NEW:
                    self.future = self.executor.submit(self._run_async_loop, bagua_ddp)
                    self.status = _AsyncInternalState.SCHEDULED
                if self.status == _AsyncInternalState.SCHEDULED:
                    with self.cv:
                        self.notified = True
                        self.cv.notify()
                    self.status = _AsyncInternalState.STARTED
                self._lock_model(bagua_ddp)

        return hook

    def init_backward_hook(self, bagua_ddp):
        def hook(parameter_name, parameter):
            if self.step_id <= self.warmup_steps and parameter_name in self._communication_tensor_names:
                bagua_ddp.mark_tensor_ready(parameter)

        return hook

    def init_post_backward_hook(self, bagua_ddp):
        def hook():
            if self.step_id <= self.warmup_steps:
                bagua_ddp.wait_pending_comm_ops()
            elif not getattr(self, "_release_after_step", False):
                self._unlock_model(bagua_ddp)

        return hook

    def need_reset(self) -> bool:
        self.step_id += 1
        if self.warmup_steps > 0 and self.step_id == self.warmup_steps + 1:
            logger.info("Async model average starts from step %d", self.step_id)
            return True
        return False

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

    def _lock_model(self, bagua_ddp):
        if not self._gated(bagua_ddp):
            self._sync_compute_stream()
        for bucket in bagua_ddp.bagua_buckets:
            if hasattr(bucket, "_async_op"):
                bucket._async_op.lock_weight()

    def _unlock_model(self, bagua_ddp):
        if not self._gated(bagua_ddp):
            self._sync_compute_stream()
        for bucket in bagua_ddp.bagua_buckets:
            if hasattr(bucket, "_async_op"):
                bucket._async_op.unlock_weight()

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
        b = bagua_ddp.bagua_buckets[0]
        return hasattr(b, "_async_op") and b._async_op.get_status()

    def _run_async_loop(self, bagua_ddp):
        with self.cv:
            while not self.notified:
                self.cv.wait()
        if torch.cuda.is_available() and comm_mod._use_cuda():
            torch.cuda.set_device(comm_mod._device_index())
        comm_step = 0
        while self._check_op_status(bagua_ddp):
            start = time.time()
            for bucket in bagua_ddp.bagua_buckets:
                for tensor in bucket.tensors:
                    tensor.bagua_mark_communication_ready_without_synchronization()
            bagua_ddp._bagua_backend.wait_pending_comm_ops(0, True)
            logger.debug("async communication cost %.2f ms, comm_step=%d", (time.time() - start) * 1000, comm_step)
            comm_step += 1
            time.sleep(self.sync_interval_ms / 1000)

    def abort(self, bagua_ddp):
        """Stop the background averaging on every rank (call after training / before evaluation)."""
        bagua_ddp = _unwrap_ddp(bagua_ddp)
        if self.status in (_AsyncInternalState.SCHEDULED, _AsyncInternalState.STARTED):
            self._unlock_model(bagua_ddp)  # a forward without backward must not keep the averaging round (and this abort) waiting
            comm_mod.barrier(comm=self.process_group.get_global_communicator())
            if hasattr(bagua_ddp.bagua_buckets[0], "_async_op"):
                bagua_ddp.bagua_buckets[0]._async_op.abort()
            with self.cv:
                self.notified = True
                self.cv.notify()
            wait([self.future])
            self.status = _AsyncInternalState.STOPPED
            logger.debug("async communication aborted.")

    def resume(self, bagua_ddp):
        """Restart the background averaging stopped by :meth:`abort` (call before training)."""
        bagua_ddp = _unwrap_ddp(bagua_ddp)
        if self.status in (_AsyncInternalState.NEW, _AsyncInternalState.STOPPED):
            comm_mod.barrier(comm=self.process_group.get_global_communicator())
            if hasattr(bagua_ddp.bagua_buckets[0], "_async_op"):
                bagua_ddp.bagua_buckets[0]._async_op.reset()
            self.notified = False
            self.future = self.executor.submit(self._run_async_loop, bagua_ddp)
            self.status = _AsyncInternalState.SCHEDULED
            logger.debug("async communication resumed.")


class AsyncModelAverageAlgorithm(Algorithm):
    def __init__(self, peer_selection_mode: str = "all", sync_interval_ms: int = 500, warmup_steps: int = 0):
        """
        Args:
            peer_selection_mode: only ``"all"`` (every worker's weights are averaged each round).
            sync_interval_ms: milliseconds between two averaging rounds.
            warmup_steps: gradient-allreduce steps before asynchronous averaging starts (0 disables the warm-up).
        """
        self.peer_selection_mode = peer_selection_mode
        self.sync_interval_ms = sync_interval_ms
        self.warmup_steps = warmup_steps

    def reify(self, process_group) -> AsyncModelAverageAlgorithmImpl:
        return AsyncModelAverageAlgorithmImpl(
            process_group, peer_selection_mode=self.peer_selection_mode, sync_interval_ms=self.sync_interval_ms, warmup_steps=self.warmup_steps
        )
