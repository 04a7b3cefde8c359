"""Per-model autotune state: history of (iteration, hyperparameters, score), the optimiser, and the bucketing rule
(reference: bagua/service/autotune_task_manager.py:1-185).

Search space: ``bucket_size = 2**k`` for k in [10, 31] × hierarchical ∈ {0,1} as in the reference, extended with the
NVSwitch allreduce kernel variant (auto / two_shot / multimem)."""
from __future__ import annotations

import collections
import csv
import math
import tempfile
from typing import Dict, List, Tuple

from ..define import BaguaHyperparameter, TensorDeclaration, TensorDtype
from .bayesian_optimizer import BayesianOptimizer, BoolParam, IntParam

__all__ = ["AutotuneTaskManager", "split_bucket_by_bucket_size"]

_UNIT = {TensorDtype.BF16.value: 2, TensorDtype.F16.value: 2, TensorDtype.F32.value: 4, TensorDtype.I64.value: 8, TensorDtype.U8.value: 1}
VARIANTS = ["auto", "two_shot", "multimem"]


def _dtype_value(d) -> str:
    return d.value if isinstance(d, TensorDtype) else str(d)


def split_bucket_by_bucket_size(tensor_list: List[TensorDeclaration], bucket_size: int, param_group_info: Dict[str, int] = {}) -> List[List[TensorDeclaration]]:
    """Greedy bucketing, dtype by dtype: append tensors until the running size reaches ``bucket_size``, then close the
    bucket (so buckets overshoot, and a bucket never mixes dtypes unless it is the trailing remainder — which is closed at
    each dtype boundary here, see below) (reference autotune_task_manager.py:85-119)."""
    buckets: List[List[TensorDeclaration]] = []
    for dtype in sorted(_UNIT.keys()):
        unit = _UNIT[dtype]
        cur: List[TensorDeclaration] = []
        cur_size = 0
        for td in tensor_list:
            if _dtype_value(td["dtype"]) != dtype:
                continue
            cur.append(td)
            cur_size += td["num_elements"] * unit
            if cur_size >= bucket_size:
                buckets.append(cur)
                cur, cur_size = [], 0
        if cur:
            # the native scheduler rejects mixed-dtype buckets, so the remainder is closed per dtype
            buckets.append(cur)
    for i in range(len(buckets)):
        buckets[i] = sorted(buckets[i], key=lambda p: param_group_info.get(p["name"], -1))
    return buckets


class AutotuneTaskManager:
    """Search loop of one model (reference autotune_task_manager.py:21-185): records (hyper-parameters, score) samples, asks the Bayesian
    optimizer for the next bucket size / hierarchical flag / kernel variant and turns it into buckets over the reported tensor order."""
    RECORD_MAX_NUM = 1000

    def __init__(self, task_name: str, need_to_log: bool) -> None:
        self.task_name = task_name
        self.record_deque = collections.deque([(-1, BaguaHyperparameter(), float("-inf"))])
        self.autotune_logfile_path = None
        if need_to_log:
            import os

            fixed = os.environ.get("BAGUA_AUTOTUNE_LOG_FILE")   # CSV of (hyper-parameters tried, score); default: a temp file
            if fixed:
                self.autotune_logfile_path = fixed
            else:
                f = tempfile.NamedTemporaryFile(prefix="bagua_autotune_", mode="w", suffix=".log", delete=False)
                self.autotune_logfile_path = f.name
                f.close()
        self.bayesian_optimizer = BayesianOptimizer(
            {
                "bucket_size_2p": IntParam(val=13, space_dimension=(10, 31)),  # 1 KiB … 2 GiB
                "is_hierarchical_reduce": BoolParam(False),
                "variant_id": IntParam(val=0, space_dimension=(0, len(VARIANTS) - 1)),
            }
        )

    # kept as a static method for parity with the reference's call sites
    split_bucket_by_bucket_size = staticmethod(split_bucket_by_bucket_size)

    @staticmethod
    def record_autotune_log(autotune_logfile_path: str, autotune_hp: dict, train_iter: int, score: float):
        cols = dict(autotune_hp)
        cols.update({"train_iter": train_iter, "score": score})
        with open(autotune_logfile_path, "a", newline="") as f:
            w = csv.DictWriter(f, fieldnames=sorted(cols.keys()))
            if f.tell() == 0:
                w.writeheader()
            w.writerow(cols)

    def tail_record(self) -> Tuple[int, BaguaHyperparameter, float]:
        return self.record_deque[-1]

    def best_hyperparameter(self) -> BaguaHyperparameter:
        return max(self.record_deque, key=lambda rec: rec[2])[1]

    def report_metrics(self, train_iter: int, hyperparameter: BaguaHyperparameter, system_efficiency_score: float) -> None:
        while len(self.record_deque) > self.RECORD_MAX_NUM:
            self.record_deque.popleft()
        self.record_deque.append((train_iter, hyperparameter, system_efficiency_score))

    def ask_hyperparmeter(self, train_iter: int, tensor_partial_order: Dict[str, int] = {}) -> BaguaHyperparameter:
        _, hp, score = self.tail_record()
        observed = {
            "bucket_size_2p": int(math.log2(max(hp.bucket_size, 1024))),
            "is_hierarchical_reduce": hp.is_hierarchical_reduce,
            "variant_id": VARIANTS.index(hp.allreduce_variant) if hp.allreduce_variant in VARIANTS else 0,
        }
        self.bayesian_optimizer.tell(observed, score)
        rec = self.bayesian_optimizer.ask()
        bucket_size = 2 ** int(rec["bucket_size_2p"])
        if self.autotune_logfile_path:
            self.record_autotune_log(self.autotune_logfile_path, observed, train_iter, score)
        tensor_list = [td for bucket in hp.buckets for td in bucket]
        tensor_list = sorted(tensor_list, key=lambda td: tensor_partial_order.get(td["name"], -1))
        return BaguaHyperparameter(
            buckets=split_bucket_by_bucket_size(tensor_list, bucket_size),
            bucket_size=bucket_size,
            is_hierarchical_reduce=bool(rec["is_hierarchical_reduce"]),
            allreduce_variant=VARIANTS[int(rec["variant_id"])],
        )
