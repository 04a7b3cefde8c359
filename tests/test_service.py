"""Autotune service (reference strategy: tests/service/test_autotune_service.py — mock workers against the REST API with a
synthetic convex score) and the Bayesian optimiser (tests/service/test_bayesian_optimizer.py)."""
import math
import threading

import numpy as np

from bagua_b200.define import BaguaHyperparameter, TensorDeclaration, TensorDtype, get_tensor_declaration_bytes
from bagua_b200.env import find_free_network_port
from bagua_b200.service import AutotuneClient, AutotuneService
from bagua_b200.service.autotune_task_manager import AutotuneTaskManager, split_bucket_by_bucket_size
from bagua_b200.service.bayesian_optimizer import BayesianOptimizer, BoolParam, FloatParam, IntParam


def test_bayesian_optimizer_finds_a_good_point():
    opt = BayesianOptimizer({"x": IntParam(5, (0, 20)), "y": FloatParam(0.0, (-1.0, 1.0)), "b": BoolParam(False)}, n_initial_points=10)

    def score(p):
        return 1.0 - ((p["x"] - 13) / 20.0) ** 2 - 0.3 * p["y"] ** 2 + (0.05 if p["b"] else 0.0)

    best = -1e9
    for _ in range(40):
        p = opt.ask()
        s = score(p)
        opt.tell(p, s)
        best = max(best, s)
    assert best > 0.95


def test_split_bucket_by_bucket_size():
    tl = [TensorDeclaration(name=f"t{i}", num_elements=1000, dtype=TensorDtype.F32) for i in range(10)]
    tl += [TensorDeclaration(name=f"h{i}", num_elements=1000, dtype=TensorDtype.BF16) for i in range(3)]
    buckets = split_bucket_by_bucket_size(tl, 8000)
    # a bucket closes once it reaches the size (so it may overshoot) and never mixes dtypes
    for b in buckets:
        assert len({td["dtype"] for td in b}) == 1
    f32 = [b for b in buckets if b[0]["dtype"] == TensorDtype.F32]
    assert [len(b) for b in f32] == [2, 2, 2, 2, 2]
    assert sum(len(b) for b in buckets) == 13
    assert get_tensor_declaration_bytes(tl[0]) == 4000 and get_tensor_declaration_bytes(tl[-1]) == 2000


def test_hyperparameter_update_roundtrip():
    hp = BaguaHyperparameter()
    hp.update({"buckets": [[{"name": "a", "num_elements": 3, "dtype": "f32"}]], "bucket_size": 123, "is_hierarchical_reduce": True, "bogus": 1})
    assert hp.bucket_size == 123 and hp.is_hierarchical_reduce and hp.buckets[0][0]["name"] == "a"
    assert BaguaHyperparameter().update(hp.dict()).dict() == hp.dict()


def test_autotune_service_with_mock_workers():
    world = 2
    service = AutotuneService(world_size=world, autotune_level=1, max_samples=12, sampling_confidence_time_s=0.0, warmup_time_s=0.0)
    port = find_free_network_port()
    server = service.make_server("127.0.0.1", port)
    threading.Thread(target=server.serve_forever, daemon=True).start()
    tensors = [TensorDeclaration(name=f"p{i}", num_elements=256 * 1024, dtype=TensorDtype.F32) for i in range(32)]  # 32 MiB

    def speed(hp: dict) -> float:  # convex score peaking at 4 MiB buckets
        k = math.log2(max(hp["bucket_size"], 1))
        return 10.0 - (k - 22) ** 2 * 0.05

    results = {}

    def worker(rank):
        c = AutotuneClient("127.0.0.1", port)
        assert c.health_check()
        assert c.register_tensors("m", tensors).status_code == 200
        it, hp, done = 0, None, False
        while not done and it < 4000:
            it += 100
            rsp = c.ask_hyperparameters("m", rank, it).json()
            hp, done = rsp["recommended_hyperparameters"], rsp["is_autotune_completed"]
            assert sum(len(b) for b in hp["buckets"]) == len(tensors)
            c.report_metrics("m", rank, it, hp, speed(hp))
        results[rank] = (hp, done, it)

    c0 = AutotuneClient("127.0.0.1", port)
    c0.report_tensor_execution_order([{"trace_id": 0, "action": "tensor_ready", "tensor_name": f"p{31 - i}", "start_time": i, "end_time": i} for i in range(32)])
    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    server.shutdown()
    assert results[0][1] and results[1][1], "autotune must complete"
    assert results[0][0]["bucket_size"] == results[1][0]["bucket_size"]
    # the best sampled configuration is handed out at the end
    mgr = service.model_dict["m"].inner
    best = mgr.best_hyperparameter()
    assert results[0][0]["bucket_size"] == best.bucket_size
    # tensors inside the final buckets follow the reported ready order (p31 first)
    flat = [td["name"] for b in results[0][0]["buckets"] for td in b]
    assert flat.index("p31") < flat.index("p0")


def test_system_autotune_measures_with_bagua_sys_perf():
    """The offline tuner end to end on CPU: a real ``bagua_sys_perf`` launch (MNIST net, 2 ranks) scored from their
    Horovod-style output (reference autotune_system.py:60-62,92-169)."""
    import os

    from bagua_b200.service import autotune_system
    from tests.mp_utils import free_port

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    old = dict(os.environ)
    os.environ.update(PYTHONPATH=repo + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="", BAGUA_FORCE_CPU="1")
    try:
        extra = ["--cpu", "--num-iters", "1", "--num-batches-per-iter", "2", "--num-warmup-batches", "1", "--batch-size", "8"]
        env, best = autotune_system.autotune_system_hyperparameters("", 2, 22, max_samples=1, model="mnist", extra_args=extra, port_fn=free_port)
        assert best > 0 and isinstance(env, dict)      # one real launch, scored from the "Total img/sec" line
    finally:
        os.environ.clear()
        os.environ.update(old)


def test_service_picks_the_kernel_variant_per_bucket_from_the_measured_table():
    """The workers hand the service the allreduce table they measured on their fabric; every bucket of every recommendation then
    carries the variant / CTA count that was fastest at the calibrated size nearest (log scale) to the bucket's own bytes."""
    world = 1
    service = AutotuneService(world_size=world, autotune_level=1, max_samples=4, sampling_confidence_time_s=0.0, warmup_time_s=0.0, default_bucket_size=1 << 20)
    port = find_free_network_port()
    server = service.make_server("127.0.0.1", port)
    threading.Thread(target=server.serve_forever, daemon=True).start()
    client = AutotuneClient("127.0.0.1", port)
    table = [{"bytes": 64 * 1024, "variant": "one_shot", "blocks": 4, "ms": 0.01, "busbw_GBs": 10.0},
             {"bytes": 1 << 20, "variant": "multimem", "blocks": 8, "ms": 0.02, "busbw_GBs": 90.0},
             {"bytes": 128 << 20, "variant": "two_shot", "blocks": 32, "ms": 0.3, "busbw_GBs": 700.0}]
    tensors = [TensorDeclaration(name="small", num_elements=8 * 1024, dtype=TensorDtype.BF16),          # 16 KiB  -> one_shot
               TensorDeclaration(name="mid", num_elements=1 << 20, dtype=TensorDtype.F32),              # 4 MiB   -> multimem
               TensorDeclaration(name="big", num_elements=40 << 20, dtype=TensorDtype.F32)]             # 160 MiB -> two_shot
    rsp = client.register_tensors("m", tensors, variant_table=table)
    hp = rsp.json()["recommended_hyperparameters"]
    by_first = {b[0]["name"]: (v, n) for b, v, n in zip(hp["buckets"], hp["bucket_variants"], hp["bucket_blocks"])}
    assert by_first["small"] == ("one_shot", 4) and by_first["mid"] == ("multimem", 8) and by_first["big"] == ("two_shot", 32), by_first
    # later recommendations (new bucket sizes from the optimiser) are annotated the same way
    for it in (100, 200, 300):
        client.report_metrics("m", 0, it, hp, speed=1.0)
        hp = client.ask_hyperparameters("m", 0, it).json()["recommended_hyperparameters"]
        assert len(hp["bucket_variants"]) == len(hp["buckets"]) or hp["allreduce_variant"] != "auto"
    # a global variant chosen by the search wins over the table
    mgr = service.model_dict["m"]
    forced = BaguaHyperparameter().update(dict(hp, allreduce_variant="two_shot"))
    assert service.apply_variant_table(mgr, forced).bucket_variants == []
    server.shutdown()


def test_setup_app_registers_the_rest_routes_on_a_flask_like_application(monkeypatch):
    """``AutotuneService.setup_app(app)`` (reference autotune_service.py:278-410): the five routes, served through Flask's calling
    convention — checked with a stand-in ``flask`` module (Flask is not installed here) against the same service object."""
    import json
    import sys
    import types

    from bagua_b200.service.autotune_service import AutotuneService

    class FakeApp:
        def __init__(self):
            self.views = {}

        def route(self, rule, methods):
            def deco(fn):
                self.views[(rule, tuple(methods))] = fn
                return fn
            return deco

    fake = types.ModuleType("flask")
    fake.request = types.SimpleNamespace(get_data=lambda as_text=True: fake.body)
    monkeypatch.setitem(sys.modules, "flask", fake)
    service = AutotuneService(world_size=1, autotune_level=0)
    app = service.setup_app(FakeApp())
    assert sorted(r for r, _ in app.views) == ["/api/v1/ask_hyperparameters", "/api/v1/health_check", "/api/v1/register_tensors", "/api/v1/report_metrics",
                                               "/api/v1/report_tensor_execution_order"]
    body, code, headers = app.views[("/api/v1/health_check", ("GET",))]()
    assert code == 200 and json.loads(body) == {"status": "ok"} and headers["Content-Type"] == "application/json"
    fake.body = json.dumps({"model_name": "m", "whether_to_bucket": True,
                            "tensor_list": [{"name": "a", "num_elements": 10, "dtype": "f32"}, {"name": "b", "num_elements": 20, "dtype": "f32"}]})
    body, code, _ = app.views[("/api/v1/register_tensors", ("POST",))]()
    reply = json.loads(body)
    assert code == 200 and [t["name"] for b in reply["recommended_hyperparameters"]["buckets"] for t in b] == ["a", "b"]
    fake.body = "{not json"
    _, code, _ = app.views[("/api/v1/report_metrics", ("POST",))]()
    assert code == 500
