"""QAdam: Adam whose first moment is communicated 8-bit-compressed after a warm-up
(reference: bagua/torch_api/algorithms/q_adam.py:1-267).

Stage 1 (``step < warmup_steps``): plain gradient allreduce + Adam.  Stage 2: the second moment is frozen, every rank
updates its local first moment ``m = β1·m + (1-β1)·g`` inside the bucket's communication program and the *momentum*
goes through the fused MinMaxUInt8 allreduce; ``need_reset()`` flips the tensor registration at the boundary."""
from __future__ import annotations

import math
from typing import List

import torch

from ...bucket import BaguaBucket
from .base import Algorithm, AlgorithmImpl

__all__ = ["QAdamOptimizer", "QAdamAlgorithm", "QAdamAlgorithmImpl"]


class QAdamOptimizer(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, warmup_steps: int = 100, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        """
        Args:
            params: parameters or param groups.
            lr: learning rate.
            warmup_steps: number of full-precision Adam steps before the moments freeze / compression starts.
            betas: coefficients of the running averages of the gradient and its square.
            eps: term added to the denominator.
            weight_decay: L2 penalty.
        """
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if warmup_steps <= 0:
            raise ValueError(f"Invalid warmup_steps parameter, must be larger than 0: {warmup_steps}")
        super().__init__(params, dict(lr=lr, warmup_steps=warmup_steps, betas=betas, eps=eps, weight_decay=weight_decay))
        self.params_in_group = []
        self.exp_avgs_in_group = []
        self.step_id = 0
        self.warmup_steps = warmup_steps
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["step"] = 0

    def __setstate__(self, state):
        super().__setstate__(state)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            lr, wd, eps = group["lr"], group["weight_decay"], group["eps"]
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step_id = st["step"]
                if p.grad is None:
                    continue
                grad = p.grad
                if wd != 0:
                    grad = grad.add(p, alpha=wd)
                if step_id < self.warmup_steps:
                    st["exp_avg"].mul_(beta1).add_(grad, alpha=1 - beta1)
                    st["exp_avg_sq"].mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                bc1 = 1 - beta1 ** step_id
                bc2 = 1 - beta2 ** step_id
                denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(eps)
                p.addcdiv_(st["exp_avg"], denom, value=-lr / bc1)
        return loss


class QAdamAlgorithmImpl(AlgorithmImpl):
    def __init__(self, process_group, q_adam_optimizer: QAdamOptimizer, hierarchical: bool = True):
        super().__init__(process_group)
        self.hierarchical = hierarchical
        self.optimizer = q_adam_optimizer
        self.warmup_steps = self.optimizer.warmup_steps

    @property
    def optimizer_step_id(self) -> int:
        param = self.optimizer.param_groups[0]["params"][0]
        return self.optimizer.state[param].get("step", 0)

    def need_reset(self) -> bool:
        if self.optimizer_step_id == self.warmup_steps:
            print(f"QAdam starts to compress from step {self.optimizer_step_id}")
            return True
        return False

    def _compressing(self) -> bool:
        return self.optimizer_step_id >= self.warmup_steps

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        parameters = bagua_ddp.bagua_build_params()
        for idx, (name, param) in enumerate(reversed(parameters)):
            param._q_adam_name = name
            param._q_adam_idx = idx
        registered = []
        compressing = self._compressing()
        for group in self.optimizer.param_groups:
            for param in group["params"]:
                if not hasattr(param, "_q_adam_name"):
                    continue
                if not compressing:
                    t = param.bagua_ensure_grad().ensure_bagua_tensor(
                        param._q_adam_name,
                        bagua_ddp.bagua_module_name,
                        getter_closure=lambda p: p.grad,
                        setter_closure=lambda p, t: setattr(p, "grad", t),
                    )
                else:
                    def set_momentum_fn(p, t):
                        self.optimizer.state[p]["exp_avg"] = t

                    t = param.bagua_ensure_grad().ensure_bagua_tensor(
                        param._q_adam_name,
                        bagua_ddp.bagua_module_name,
                        getter_closure=lambda p: self.optimizer.state[p]["exp_avg"],
                        setter_closure=set_momentum_fn,
                    )
                registered.append(t)
        registered.sort(key=lambda p: p._q_adam_idx)
        self._communication_tensor_names = set(p._q_adam_name for p in registered)
        return registered

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        n = self.process_group.size()
        if self._compressing():
            from .bytegrad import bytegrad_min_bucket_bytes, merge_small_buckets

            tensors = merge_small_buckets(tensors, bytegrad_min_bucket_bytes(tensors))   # fixed cost per quantised exchange: see bytegrad.py
        return [BaguaBucket(b, flatten=do_flatten, name=str(i), alignment=32 * n, group=self.process_group) for i, b in enumerate(tensors)]

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        bucket.clear_ops()
        if not self._compressing():
            bucket.append_centralized_synchronous_op(hierarchical=False, average=True, group=self.process_group)
            return

        beta1, _b2 = self.optimizer.param_groups[0]["betas"]
        flat = bucket.backend_tensor
        if flat is not None and flat.is_cuda and bucket._engine(self.process_group) is not None:
            # NVSwitch path: the momentum update m = β1·m + (1−β1)·g runs INSIDE the fused ByteGrad kernel's first pass (the
            # reference needs a python op on the comm thread here, q_adam.py:193-221). For that the gradients get the same flat
            # layout as the momentum bucket: every p.grad becomes a view of one mirror buffer (zero_grad keeps the views).
            from ...tensor import dense_strides

            gflat = torch.zeros_like(flat)
            base, es = flat.data_ptr(), flat.element_size()
            with torch.no_grad():
                for t in bucket.tensors:
                    m = t.bagua_getter_closure()
                    view = torch.as_strided(gflat, t.shape, dense_strides(m), (m.data_ptr() - base) // es)
                    if t.grad is not None:
                        view.copy_(t.grad)
                    t.grad = view
            bucket._qadam_grad_flat = gflat
            bucket.append_centralized_synchronous_op(hierarchical=self.hierarchical, average=True, scattergather=True, compression="MinMaxUInt8",
                                                     group=self.process_group, momentum_source=(gflat, beta1))
            return

        def calculate_momentum(*_):
            moms = [t.bagua_getter_closure() for t in bucket.tensors]
            grads = [t.grad for t in bucket.tensors]
            with torch.no_grad():
                # m = β1·m + (1-β1)·g for the whole bucket in one multi-tensor launch
                torch._foreach_lerp_(moms, grads, 1 - beta1)

        bucket.append_python_op(calculate_momentum, group=self.process_group)
        bucket.append_centralized_synchronous_op(
            hierarchical=self.hierarchical, average=True, scattergather=True, compression="MinMaxUInt8", group=self.process_group
        )

    def init_backward_hook(self, bagua_ddp):
        compressing = self._compressing()
        names = self._communication_tensor_names

        def hook(parameter_name, parameter):
            if parameter_name not in names:
                return
            eff = self.optimizer.state[parameter]["exp_avg"] if compressing else parameter.grad
            assert parameter._bagua_backend_tensor.data_ptr() == eff.data_ptr(), (
                "bagua backend tensor data_ptr should match " + ("momentum data_ptr in QAdam compression stage" if compressing else "grad data_ptr in QAdam warm-up stage")
            )
            bagua_ddp.mark_tensor_ready(parameter)

        return hook


class QAdamAlgorithm(Algorithm):
    def __init__(self, q_adam_optimizer: QAdamOptimizer, hierarchical: bool = True):
        self.hierarchical = hierarchical
        self.optimizer = q_adam_optimizer

    def reify(self, process_group) -> QAdamAlgorithmImpl:
        return QAdamAlgorithmImpl(process_group, q_adam_optimizer=self.optimizer, hierarchical=self.hierarchical)
