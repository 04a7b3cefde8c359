"""Centralized synchronous full-precision data parallelism (reference: bagua/torch_api/algorithms/gradient_allreduce.py:1-64).

On NVSwitch each bucket is reduced by ONE kernel (two-shot peer loads/stores or NVLS ``multimem``) that also applies
the 1/N average; ``fused_optimizer=True`` additionally folds the SGD update of the rank's 1/N shard and the all-gather
of the updated weights into that kernel (see ``FusedShardedSGD``)."""
from __future__ import annotations

from .base import Algorithm, AlgorithmImpl

__all__ = ["GradientAllReduceAlgorithm", "GradientAllReduceAlgorithmImpl"]


class GradientAllReduceAlgorithmImpl(AlgorithmImpl):
    def __init__(self, process_group, hierarchical: bool = False, average: bool = True, variant: str = "auto"):
        super().__init__(process_group)
        self.hierarchical = hierarchical
        self.average = average
        self.variant = variant

    def init_operations(self, bagua_ddp, bucket):
        bucket.clear_ops()
        hp = getattr(bagua_ddp, "_bagua_hyperparameters", None)
        variant = self.variant
        if variant == "auto" and hp is not None and getattr(hp, "allreduce_variant", "auto") != "auto":
            variant = hp.allreduce_variant
        bucket.append_centralized_synchronous_op(hierarchical=self.hierarchical, average=self.average, group=self.process_group, variant=variant)


class GradientAllReduceAlgorithm(Algorithm):
    def __init__(self, hierarchical: bool = False, average: bool = True, variant: str = "auto"):
        """
        Args:
            hierarchical: intra-node reduce → inter-node allreduce → intra-node broadcast when the job spans nodes.
            average: average (``True``) or sum gradients.
            variant: ``auto`` | ``one_shot`` | ``two_shot`` | ``multimem`` kernel choice (``auto`` = by size / autotune).
        """
        self.hierarchical = hierarchical
        self.average = average
        self.variant = variant

    def reify(self, process_group) -> GradientAllReduceAlgorithmImpl:
        return GradientAllReduceAlgorithmImpl(process_group, hierarchical=self.hierarchical, average=self.average, variant=self.variant)


# ---------------------------------------------------------------------------------------------------------------
# Gradient allreduce with the optimizer folded into the communication kernel
# ---------------------------------------------------------------------------------------------------------------
class ShardedFusedSGD(object):
    """Marker mix-in: see :func:`make_sharded_fused_sgd`."""


def make_sharded_fused_sgd(params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
    """SGD whose update runs *inside* each bucket's NVSwitch kernel (``allreduce_sgd_kernel``): reduce-scatter of the
    gradients → SGD(momentum) on this rank's 1/N shard with fp32 master weights → all-gather of the updated weights, one
    launch per bucket, overlapped with the rest of backward.  Optimizer state is sharded N ways (ZeRO-1-like) and the
    gradient bucket is cleared on the way out.  Use together with ``FusedGradientAllReduceAlgorithm``; with one process (or
    without a peer engine) it behaves like :class:`bagua_b200.ops.optim.FusedSGD`."""
    from ...ops.optim import FusedSGD

    class _ShardedFusedSGD(FusedSGD, ShardedFusedSGD):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self._comm_ops = []   # AllReduceSgdOp per bucket once the algorithm has taken over

        def _sync_hyper(self):
            g = self.param_groups[0]
            for op in self._comm_ops:
                op.set_hyper(float(g["lr"]), float(g["momentum"]), float(g["dampening"]), float(g["weight_decay"]), bool(g["nesterov"]))

        def step(self, closure=None):
            if not self._comm_ops:
                return super().step(closure)
            # the update already happened in the bucket kernels of this iteration; publish hyper-parameters for the next
            self._sync_hyper()
            self._grads_zeroed = True
            self.kernel_launches += len(self._comm_ops)
            return None

    return _ShardedFusedSGD(params, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov,
                            master_weights=True, zero_grad_in_step=True)


class ShardedFusedAdam(object):
    """Marker mix-in: see :func:`make_sharded_fused_adam`."""


def make_sharded_fused_adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, adamw=False):
    """Adam / AdamW counterpart of :func:`make_sharded_fused_sgd`: the update of this rank's 1/N shard (fp32 master weights
    and both moments) runs inside each bucket's reduce-scatter → all-gather kernel (``allreduce_adam_kernel``).  With one
    process (or without a peer engine) it behaves like :class:`bagua_b200.ops.optim.FusedAdam`."""
    from ...ops.optim import FusedAdam

    class _ShardedFusedAdam(FusedAdam, ShardedFusedAdam):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self._comm_ops = []

        def _sync_hyper(self):
            g = self.param_groups[0]
            for op in self._comm_ops:
                op.set_hyper(float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), bool(g["adamw"]))

        def step(self, closure=None):
            if not self._comm_ops:
                return super().step(closure)
            self._sync_hyper()
            self._grads_zeroed = True
            self.kernel_launches += len(self._comm_ops)
            return None

    return _ShardedFusedAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adamw=adamw, master_weights=True,
                             zero_grad_in_step=True)


class FusedGradientAllReduceAlgorithmImpl(GradientAllReduceAlgorithmImpl):
    def __init__(self, process_group, optimizer, average: bool = True):
        super().__init__(process_group, hierarchical=False, average=average)
        self.optimizer = optimizer
        self._weight_slices = []

    def init_forward_pre_hook(self, bagua_ddp):
        """The update of iteration *i* runs inside the bucket kernels of *i*'s backward pass, before ``optimizer.step()`` is
        called — so the hyper-parameters are published at the start of every iteration (after any ``lr_scheduler.step()`` of
        the previous one), not only from ``step()``."""
        opt = self.optimizer

        def hook(input):
            if getattr(opt, "_comm_ops", None):
                opt._sync_hyper()

        return hook

    def tensors_to_buckets(self, tensors, do_flatten):
        from ...bucket import BaguaBucket

        assert do_flatten, "the fused optimizer needs flattened buckets"
        es = tensors[0][0].bagua_getter_closure().element_size()
        return [BaguaBucket(b, flatten=True, name=str(i), alignment=max(1, 16 // es), group=self.process_group) for i, b in enumerate(tensors)]

    def init_operations(self, bagua_ddp, bucket):
        import os

        import torch

        from ...core import dtype_code, native
        from ...tensor import dense_strides

        bucket.clear_ops()
        eng = bucket._engine(self.process_group)
        opt = self.optimizer
        is_adam = isinstance(opt, ShardedFusedAdam)
        if eng is None or bucket._slice is None or not (isinstance(opt, ShardedFusedSGD) or is_adam):
            return super().init_operations(bagua_ddp, bucket)
        assert len(opt.param_groups) == 1, "the in-kernel optimizer supports a single parameter group"
        C = native()
        flat = bucket.backend_tensor
        n, rank = self.process_group.size(), self.process_group.rank()
        nbytes = flat.numel() * flat.element_size()
        assert nbytes % 16 == 0
        # weights mirror the gradient bucket layout inside their own symmetric slice (peers write them in the kernel)
        wslice = eng.alloc(nbytes)
        bucket._companion_slices.append(wslice)
        wflat = wslice.view(flat.dtype, flat.numel())
        wflat.zero_()
        base = flat.data_ptr()
        with torch.no_grad():
            for t in bucket.tensors:
                g = t.bagua_getter_closure()
                assert t.dtype == flat.dtype, "weights and gradients must share a dtype for the fused kernel"
                off = (g.data_ptr() - base) // flat.element_size()
                view = torch.as_strided(wflat, t.shape, dense_strides(t), off)
                view.copy_(t.data)
                t.data = view
        vecs = nbytes // 16
        vpr = (vecs + n - 1) // n
        per = 16 // flat.element_size()
        lo, hi = rank * vpr * per, min((rank + 1) * vpr * per, flat.numel())
        master = torch.zeros(vpr * per, dtype=torch.float32, device=flat.device)
        if hi > lo:
            master[: hi - lo].copy_(wflat[lo:hi].float())
        momentum = torch.zeros(vpr * per, dtype=torch.float32, device=flat.device)
        # NVLS whenever the fabric offers it (the variant validated on 2 and 8 GPUs); peer ld/st two-shot otherwise
        use_mc = bool(wslice.has_multicast and bucket._slice.has_multicast and eng.has_multicast)
        if os.environ.get("BAGUA_FUSED_MULTIMEM", "auto") == "0":   # experiment switch: peer ld/st flavour of the fused kernel
            use_mc = False
        # 16 CTAs: the configuration measured at 56 994 img/s on 8 GPUs (profiles/bench8_fused.json); the kernel also streams
        # the fp32 optimizer shard, so it wants more CTAs than the bare multimem allreduce (8)
        cfg = eng.launch_cfg("multimem" if use_mc else "two_shot", nbytes, blocks=int(os.environ.get("BAGUA_FUSED_BLOCKS", "0")) or (16 if use_mc else 32))
        scale = (1.0 / n) if self.average else 1.0
        if is_adam:
            second = torch.zeros(vpr * per, dtype=torch.float32, device=flat.device)
            bucket._fused_state = (master, momentum, second)
            op = C.AllReduceAdamOp(eng.comm, bucket._slice.buf, wslice.buf, bucket._slice.offset, wslice.offset, nbytes, dtype_code(flat.dtype),
                                   master.data_ptr(), momentum.data_ptr(), second.data_ptr(), scale, True, use_mc, cfg)
        else:
            bucket._fused_state = (master, momentum)
            op = C.AllReduceSgdOp(eng.comm, bucket._slice.buf, wslice.buf, bucket._slice.offset, wslice.offset, nbytes, dtype_code(flat.dtype),
                                  master.data_ptr(), momentum.data_ptr(), scale, True, use_mc, cfg)
        bucket.backend_bucket.append_op(op)
        bucket._ops_keepalive.append(op)
        bucket.allreduce_variant = ("fused_adam_" if is_adam else "fused_sgd_") + ("multimem" if use_mc else "two_shot")
        opt._comm_ops.append(op)
        opt._sync_hyper()


class FusedGradientAllReduceAlgorithm(Algorithm):
    """Gradient allreduce whose buckets also apply the optimizer update (see :func:`make_sharded_fused_sgd` /
    :func:`make_sharded_fused_adam`)."""

    def __init__(self, optimizer, average: bool = True):
        self.optimizer = optimizer
        self.average = average

    def reify(self, process_group):
        self.optimizer._comm_ops = []
        return FusedGradientAllReduceAlgorithmImpl(process_group, self.optimizer, average=self.average)
