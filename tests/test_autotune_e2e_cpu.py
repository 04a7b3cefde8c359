"""Engine ↔ autotune service integration on CPU/gloo: rank 0 hosts the service process, every rank registers tensors,
reports speed every 100 iterations and re-buckets when the service hands out new hyperparameters
(reference flow: bagua_distributed.py:325-391 + autotune_service.py)."""
import torch

from tests.mp_utils import run_distributed


def _worker(rank, world):
    import torch.nn as nn
    import torch.nn.functional as F

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import gradient_allreduce

    bagua.init_process_group()
    assert bagua.communication.get_autotune_service_port() is not None
    torch.manual_seed(0)
    model = nn.Sequential(*[nn.Linear(64, 64) for _ in range(6)], nn.Linear(64, 4))
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    model = model.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
    layouts = set()
    for it in range(420):
        x, y = torch.randn(8, 64), torch.randint(0, 4, (8,))
        opt.zero_grad()
        F.cross_entropy(model(x), y).backward()
        opt.step()
        layouts.add(tuple(len(b.tensors) for b in model.bagua_buckets))
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    return flat, len(layouts), model.bagua_ddp._bagua_hyperparameters.bucket_size


def test_autotune_rebuckets_and_keeps_replicas_in_sync():
    env = {"BAGUA_AUTOTUNE": "1", "BAGUA_AUTOTUNE_WARMUP_TIME_S": "0", "BAGUA_AUTOTUNE_SAMPLING_CONFIDENCE_TIME_S": "0",
           "BAGUA_AUTOTUNE_MAX_SAMPLES": "3", "BAGUA_DEFAULT_BUCKET_SIZE": "20000"}
    res = run_distributed(_worker, world=2, extra_env=env, timeout=400)
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][1] >= 2, "the service should have proposed at least one different bucketing"
    assert res[0][2] == res[1][2]
