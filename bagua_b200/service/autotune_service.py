"""Autotune HTTP service + client (reference: bagua/service/autotune_service.py:1-460).

Same REST contract (``/api/v1/{register_tensors,report_metrics,ask_hyperparameters,report_tensor_execution_order,
health_check}``) and the same gating (all ranks at the same iteration, warm-up, confidence time, ``max_samples``) on the
python standard library's threaded HTTP server — flask/gevent are not dependencies."""
from __future__ import annotations

import copy
import json
import logging
import multiprocessing
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Dict, List, Optional

import requests

from ..define import BaguaHyperparameter, TensorDeclaration, TensorDtype
from .autotune_task_manager import AutotuneTaskManager, split_bucket_by_bucket_size

__all__ = ["AutotuneService", "AutotuneClient", "AutotuneServiceTaskManager", "start_autotune_server_process", "run_autotune_server"]

logger = logging.getLogger(__name__)


class _Encoder(json.JSONEncoder):
    def default(self, obj):
        try:
            import numpy as np

            if isinstance(obj, np.integer):
                return int(obj)
            if isinstance(obj, np.floating):
                return float(obj)
            if isinstance(obj, np.ndarray):
                return obj.tolist()
        except Exception:  # noqa: BLE001
            pass
        if isinstance(obj, TensorDtype):
            return obj.value
        return super().default(obj)


class AutotuneServiceTaskManager:
    """Per-model tuning session of the service: the task manager (search), the check board (which rank reported which iteration), the
    sampling clock and the hyper-parameters currently in force (reference autotune_service.py:35-45)."""
    def __init__(self, task_name: str, world_size: int, is_output_autotune_log: bool) -> None:
        self.inner = AutotuneTaskManager(task_name, is_output_autotune_log)
        self.warmup_pass_count = 0
        self.sampling_count = 0
        self.lock = threading.Lock()
        self.check_board = [-1] * world_size
        self.time_hp_last_granted = time.time()
        self.hyperparameter = BaguaHyperparameter()
        self.variant_table: List[dict] = []   # measured by the workers: [{"bytes", "variant", "blocks", "ms", "busbw_GBs"}, ...]


class AutotuneService:
    """The autotune HTTP service run by rank 0 (reference autotune_service.py:48-303): workers register their tensors, report speed and
    tensor ready order, and ask for the next bucketing; a new sample is only issued when every rank has reached the same iteration."""
    MAX_TRACE_INFO = 1000

    def __init__(self, world_size, autotune_level=0, max_samples=60, sampling_confidence_time_s=5, warmup_time_s=30,
                 is_output_autotune_log=False, default_bucket_size=10 * 1024 ** 2):
        self.autotune_level = autotune_level
        self.world_size = world_size
        self.max_samples = max_samples
        self.sampling_confidence_time_s = sampling_confidence_time_s
        self.warmup_time_s = warmup_time_s
        self.is_output_autotune_log = is_output_autotune_log
        self.default_bucket_size = default_bucket_size
        self.model_dict: Dict[str, AutotuneServiceTaskManager] = {}
        self.model_dict_mutex = threading.Lock()
        self.trace_info_dict = {}
        self.tensor_partial_order: Dict[str, int] = {}
        self.tensor_partial_order_lock = threading.Lock()

    # -- decision ------------------------------------------------------------------------------------------------
    def autotune(self, hp_manager: AutotuneServiceTaskManager, rank: int, train_iter: int, tensor_partial_order: Dict[str, int] = {}):
        if hp_manager.sampling_count > self.max_samples:
            return
        sampling_time = time.time() - hp_manager.time_hp_last_granted
        # skip at least once during warm-up
        if sampling_time < self.warmup_time_s or hp_manager.warmup_pass_count == 0:
            hp_manager.warmup_pass_count += 1
            return
        if hp_manager.sampling_count == 0:
            if sampling_time < self.warmup_time_s + self.sampling_confidence_time_s:
                return
        elif sampling_time < self.sampling_confidence_time_s:
            return
        recommended = hp_manager.inner.ask_hyperparmeter(train_iter, tensor_partial_order)
        if hp_manager.sampling_count < self.max_samples:
            hp_manager.hyperparameter = recommended
        else:
            hp_manager.hyperparameter = hp_manager.inner.best_hyperparameter()
        self.apply_variant_table(hp_manager, hp_manager.hyperparameter)
        hp_manager.sampling_count += 1
        hp_manager.time_hp_last_granted = time.time()

    # -- kernel variant per message size, from measured bus bandwidth ---------------------------------------------------
    @staticmethod
    def _bucket_bytes(bucket) -> int:
        unit = {"f32": 4, "f16": 2, "bf16": 2, "u8": 1, "i64": 8}
        return sum(int(td["num_elements"]) * unit.get(str(getattr(td["dtype"], "value", td["dtype"])), 4) for td in bucket)

    def apply_variant_table(self, mgr: AutotuneServiceTaskManager, hp: BaguaHyperparameter) -> BaguaHyperparameter:
        """Every bucket gets the allreduce kernel variant (and CTA count) that was fastest, at the calibrated message size nearest
        to the bucket's own (log scale), in the table the workers measured on their fabric — unless the search itself is trying
        a global variant (``allreduce_variant != "auto"``)."""
        import math

        if not mgr.variant_table or hp.allreduce_variant != "auto":
            hp.bucket_variants, hp.bucket_blocks = [], []
            return hp
        vs, bs = [], []
        for b in hp.buckets:
            n = max(self._bucket_bytes(b), 1)
            row = min(mgr.variant_table, key=lambda r: abs(math.log2(n) - math.log2(max(int(r["bytes"]), 1))))
            vs.append(str(row["variant"]))
            bs.append(int(row.get("blocks", 0)))
        hp.bucket_variants, hp.bucket_blocks = vs, bs
        return hp

    # -- endpoints -----------------------------------------------------------------------------------------------
    def register_tensors(self, req: dict):
        model_name: str = req["model_name"]
        tensor_list: List[TensorDeclaration] = req["tensor_list"]
        whether_to_bucket: bool = req.get("whether_to_bucket", True)
        with self.model_dict_mutex:
            if model_name not in self.model_dict:
                self.model_dict[model_name] = AutotuneServiceTaskManager(model_name, self.world_size, self.is_output_autotune_log)
        mgr = self.model_dict[model_name]
        bucket_size = self.default_bucket_size if whether_to_bucket else 10 * 1024 ** 5
        with mgr.lock:
            if req.get("variant_table"):
                mgr.variant_table = list(req["variant_table"])
            hp = BaguaHyperparameter(buckets=split_bucket_by_bucket_size(tensor_list, bucket_size), bucket_size=bucket_size)
            self.apply_variant_table(mgr, hp)
            mgr.time_hp_last_granted = time.time()
            mgr.hyperparameter = hp
            return 200, {"recommended_hyperparameters": hp.dict()}

    def report_metrics(self, req: dict):
        model_name = req["model_name"]
        if model_name not in self.model_dict:
            return 405, "Service not ready for report_metrics!"
        mgr = self.model_dict[model_name]
        with mgr.lock:
            last_iter, _, _ = mgr.inner.tail_record()
            if req["train_iter"] <= last_iter:  # only the first report of an iteration counts
                return 200, {}
            mgr.inner.report_metrics(
                train_iter=req["train_iter"],
                hyperparameter=BaguaHyperparameter().update(req["hyperparameters"]),
                system_efficiency_score=req["speed"],
            )
        return 200, {}

    def ask_hyperparameters(self, req: dict):
        model_name = req["model_name"]
        rank, train_iter = req["rank"], req["train_iter"]
        if model_name not in self.model_dict:
            return 405, "Service not ready for report_metrics!"
        mgr = self.model_dict[model_name]
        with self.tensor_partial_order_lock:
            order = copy.deepcopy(self.tensor_partial_order)
        with mgr.lock:
            board = mgr.check_board
            # autotune only when (1) enabled, (2) no hyperparameter roll-out in progress (all ranks at the same
            # iteration), (3) at most once per iteration
            if self.autotune_level >= 1 and board.count(board[0]) == len(board) and board[rank] < train_iter:
                self.autotune(mgr, rank, train_iter, order)
            board[rank] = train_iter
            return 200, {
                "recommended_hyperparameters": mgr.hyperparameter.dict(),
                "is_autotune_completed": mgr.sampling_count > self.max_samples,
            }

    def report_tensor_execution_order(self, req: dict):
        spans = sorted(req["spans"], key=lambda s: s["start_time"])
        with self.tensor_partial_order_lock:
            for span in spans:
                key = (span["tensor_name"], span["action"])
                if key in self.trace_info_dict:
                    continue
                self.trace_info_dict[key] = True
                if span["tensor_name"] not in self.tensor_partial_order:
                    self.tensor_partial_order[span["tensor_name"]] = len(self.tensor_partial_order)
        return 200, {}

    def health_check(self, _req=None):
        return 200, {"status": "ok"}

    # -- http plumbing -----------------------------------------------------------------------------------------------
    def make_server(self, host: str, port: int) -> ThreadingHTTPServer:
        service = self
        routes = {
            "/api/v1/register_tensors": service.register_tensors,
            "/api/v1/report_metrics": service.report_metrics,
            "/api/v1/ask_hyperparameters": service.ask_hyperparameters,
            "/api/v1/report_tensor_execution_order": service.report_tensor_execution_order,
        }

        class Handler(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, fmt, *args):  # quiet
                logger.debug("autotune http: " + fmt, *args)

            def _send(self, code: int, body):
                data = (body if isinstance(body, str) else json.dumps(body, cls=_Encoder)).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            def do_GET(self):
                if self.path == "/api/v1/health_check":
                    self._send(*service.health_check())
                else:
                    self._send(404, "not found")

            def do_POST(self):
                fn = routes.get(self.path)
                if fn is None:
                    self._send(404, "not found")
                    return
                n = int(self.headers.get("Content-Length", 0))
                try:
                    req = json.loads(self.rfile.read(n) or b"{}")
                    code, body = fn(req)
                except Exception as e:  # noqa: BLE001
                    logger.exception("autotune service error")
                    code, body = 500, str(e)
                self._send(code, body)

        server = ThreadingHTTPServer((host, port), Handler)
        server.daemon_threads = True
        return server


    def setup_app(self, app):
        """Register the same five routes on a Flask application (``app.route(rule, methods=[...])``) — how the reference builds its
        server (autotune_service.py:278-410, ``AutotuneService.setup_app``).  The built-in server of :meth:`make_server` does not
        need Flask; this is for embedding the service into an existing Flask / WSGI deployment.  Returns ``app``."""
        service = self

        def bind(rule: str, fn, methods):
            def view():
                try:
                    req = None
                    if "POST" in methods:
                        from flask import request   # only a Flask deployment gets here

                        req = json.loads(request.get_data(as_text=True) or "{}")
                    code, body = fn(req)
                except Exception as e:  # noqa: BLE001
                    logger.exception("autotune service error")
                    code, body = 500, str(e)
                return (body if isinstance(body, str) else json.dumps(body, cls=_Encoder)), code, {"Content-Type": "application/json"}

            view.__name__ = "bagua_autotune_" + rule.rsplit("/", 1)[-1]
            app.route(rule, methods=list(methods))(view)

        bind("/api/v1/register_tensors", service.register_tensors, ("POST",))
        bind("/api/v1/report_metrics", service.report_metrics, ("POST",))
        bind("/api/v1/ask_hyperparameters", service.ask_hyperparameters, ("POST",))
        bind("/api/v1/report_tensor_execution_order", service.report_tensor_execution_order, ("POST",))
        bind("/api/v1/health_check", service.health_check, ("GET",))
        return app


def run_autotune_server(port: int, world_size: int, **kwargs):
    """Serve forever (entry point of the daemon process on rank 0)."""
    from .. import env

    service = AutotuneService(
        world_size=world_size,
        autotune_level=kwargs.get("autotune_level", env.get_autotune_level()),
        max_samples=kwargs.get("max_samples", env.get_autotune_max_samples()),
        sampling_confidence_time_s=kwargs.get("sampling_confidence_time_s", env.get_autotune_sampling_confidence_time_s()),
        warmup_time_s=kwargs.get("warmup_time_s", env.get_autotune_warmup_time_s()),
        is_output_autotune_log=kwargs.get("is_output_autotune_log", env.get_is_output_autotune_log()),
        default_bucket_size=kwargs.get("default_bucket_size", env.get_default_bucket_size()),
    )
    server = service.make_server("0.0.0.0", port)
    server.serve_forever()


class _ServerProcess:
    """Handle of the service process with the small part of ``multiprocessing.Process`` its users need."""

    def __init__(self, popen):
        self._p = popen
        self.pid = popen.pid

    def is_alive(self) -> bool:
        return self._p.poll() is None

    def terminate(self):
        if self.is_alive():
            self._p.terminate()

    def kill(self):
        if self.is_alive():
            self._p.kill()

    def join(self, timeout=None):
        try:
            self._p.wait(timeout=timeout)
        except Exception:  # noqa: BLE001
            pass


def start_autotune_server_process(port: int, world_size: int, **kwargs) -> "_ServerProcess":
    """Start :class:`AutotuneService` on ``port`` in its own interpreter (``python -m bagua_b200.service.autotune_service``) and
    return a handle.  Deliberately NOT ``multiprocessing`` with the spawn start method: that re-imports the parent's ``__main__``
    — the user's training script, which has no reason to carry an ``if __name__ == "__main__"`` guard — and the child would call
    ``init_process_group`` as a second rank 0 and hang the job; fork is not an option once CUDA / NCCL are initialised."""
    import atexit
    import os
    import subprocess
    import sys

    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "CUDA_VISIBLE_DEVICES")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "bagua_b200.service.autotune_service", "--port", str(port), "--world-size", str(world_size), "--kwargs", json.dumps(kwargs)]
    p = _ServerProcess(subprocess.Popen(cmd, env=env, stdin=subprocess.DEVNULL))
    atexit.register(p.terminate)
    return p


NpEncoder = _Encoder  # name used by the reference (autotune_service.py:21-32)


def reset_error_retry(request_func, max_retries: int = 3, delay_s: float = 1.0):
    """Decorator: repeat a request when the connection is reset by the peer (reference autotune_service.py:306-322)."""
    import functools

    @functools.wraps(request_func)
    def wrapper(*args, **kwargs):
        for attempt in range(max_retries + 1):
            try:
                return request_func(*args, **kwargs)
            except (ConnectionResetError, requests.exceptions.ConnectionError) as e:
                if attempt == max_retries:
                    raise
                logging.warning("request failed (attempt %d): %s", attempt, e)
                time.sleep(delay_s)

    return wrapper


class AutotuneClient:
    """REST client with keep-alive and retries (reference autotune_service.py:306-435)."""

    def __init__(self, service_addr: str, service_port: int, proxies: Optional[dict] = None, timeout: float = 5.0):
        self.base = f"http://{service_addr}:{service_port}"
        self.health_timeout = timeout
        self.session = requests.Session()
        self.session.trust_env = False
        self.proxies = proxies or {"http": None, "https": None}
        adapter = requests.adapters.HTTPAdapter(max_retries=3)
        self.session.mount("http://", adapter)

    def _post(self, path: str, payload: dict):
        last = None
        for _ in range(3):
            try:
                return self.session.post(self.base + path, data=json.dumps(payload, cls=_Encoder), proxies=self.proxies, timeout=60)
            except requests.exceptions.ConnectionError as e:  # connection reset by peer etc.
                last = e
                time.sleep(0.1)
        raise last

    def report_metrics(self, model_name: str, rank: int, train_iter: int, hyperparameters: dict, speed: float):
        return self._post("/api/v1/report_metrics", {"model_name": model_name, "rank": rank, "train_iter": train_iter, "hyperparameters": hyperparameters, "speed": speed})

    def register_tensors(self, model_name: str, tensor_list: List[TensorDeclaration], whether_to_bucket: bool = True, variant_table: Optional[list] = None):
        payload = {"model_name": model_name, "tensor_list": tensor_list, "whether_to_bucket": whether_to_bucket}
        if variant_table:
            payload["variant_table"] = variant_table   # measured allreduce bus bandwidth per size class (PeerEngine.calibrate)
        return self._post("/api/v1/register_tensors", payload)

    def ask_hyperparameters(self, model_name: str, rank: int, train_iter: int):
        return self._post("/api/v1/ask_hyperparameters", {"model_name": model_name, "rank": rank, "train_iter": train_iter})

    def report_tensor_execution_order(self, spans: List[dict]):
        return self._post("/api/v1/report_tensor_execution_order", {"spans": spans})

    def health_check(self) -> bool:
        try:
            r = self.session.get(self.base + "/api/v1/health_check", proxies=self.proxies, timeout=self.health_timeout)
            return r.status_code == 200
        except Exception:  # noqa: BLE001
            return False


def _main():
    import argparse

    ap = argparse.ArgumentParser(description="bagua_b200 autotune service")
    ap.add_argument("--port", type=int, required=True)
    ap.add_argument("--world-size", type=int, required=True)
    ap.add_argument("--kwargs", default="{}")
    a = ap.parse_args()
    run_autotune_server(a.port, a.world_size, **json.loads(a.kwargs))


if __name__ == "__main__":
    _main()
