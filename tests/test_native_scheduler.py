"""C++ CommScheduler semantics (reference behaviours: lib.rs:270-337): strict in-order scheduling, duplicate detection,
python ops on the worker thread, waitable completion, watchdog."""
import threading
import time

import pytest

from bagua_b200.core import native


def make(names, base=0x1000):
    C = native()
    return [C.Tensor(n, base + 64 * i, 4, 0, -1) for i, n in enumerate(names)]


def test_in_order_scheduling_and_python_ops():
    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    order = []
    t0, t1 = make(["a", "b"]), make(["c"], base=0x2000)
    b0, b1 = C.Bucket("b0", t0), C.Bucket("b1", t1)
    b0.append_python_op(lambda name: order.append(name))
    b1.append_python_op(lambda name: order.append(name))
    be.register_ordered_buckets([b0, b1])
    # bucket 1 becomes ready first but must wait behind bucket 0
    be.mark_communication_ready(t1[0], 0)
    time.sleep(0.05)
    assert order == []
    be.mark_communication_ready(t0[0], 0)
    be.mark_communication_ready(t0[1], 0)
    assert be.wait_pending_comm_ops(0, True) == 2
    assert order == ["b0", "b1"]
    # a second iteration works (flags are reset, buckets rotate)
    for t in t0 + t1:
        be.mark_communication_ready(t, 0)
    assert be.wait_pending_comm_ops(0, True) == 2
    assert order == ["b0", "b1", "b0", "b1"]
    assert be.scheduled_total() == 4
    be.shutdown()


def test_duplicate_detection():
    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    a = C.Tensor("x", 0x1000, 4, 0, -1)
    b = C.Tensor("x", 0x2000, 4, 0, -1)
    with pytest.raises(ValueError):
        be.register_ordered_buckets([C.Bucket("b0", [a]), C.Bucket("b1", [b])])
    c = C.Tensor("y", 0x1000, 4, 0, -1)
    with pytest.raises(ValueError):
        be.register_ordered_buckets([C.Bucket("b0", [a]), C.Bucket("b1", [c])])
    with pytest.raises(ValueError):
        C.Bucket("mixed", [a, C.Tensor("z", 0x3000, 4, 1, -1)])  # mixed dtypes
    be.shutdown()


def test_padding_tensor_is_always_ready_and_contiguity():
    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    t = C.Tensor("w", 0x1000, 4, 0, -1)
    pad = C.Tensor("pad", 0x1010, 4, 0, -1)
    b = C.Bucket("b", [t, pad])
    assert b.contiguous() and b.flat_ptr() == 0x1000 and b.bytes() == 32
    b.mark_padding(1)
    hits = []
    b.append_python_op(lambda n: hits.append(n))
    be.register_ordered_buckets([b])
    be.mark_communication_ready(t, 0)
    assert be.wait_pending_comm_ops(0, True) == 1 and hits == ["b"]
    be.shutdown()


def test_python_op_error_is_reported():
    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    (t,) = make(["e"])
    b = C.Bucket("b", [t])

    def boom(_):
        raise RuntimeError("boom")

    b.append_python_op(boom)
    be.register_ordered_buckets([b])
    be.mark_communication_ready(t, 0)
    with pytest.raises(RuntimeError, match="boom"):
        be.wait_pending_comm_ops(0, True)
    be.shutdown()


def test_watchdog_reports_stuck_op():
    C = native()
    be = C.Backend(10, -1, 0, 0.5)
    be.set_watchdog_fatal(False)
    (t,) = make(["s"])
    b = C.Bucket("b", [t])
    release = threading.Event()
    b.append_python_op(lambda _: release.wait(5))
    be.register_ordered_buckets([b])
    be.mark_communication_ready(t, 0)
    with pytest.raises(RuntimeError, match="watchdog"):
        be.wait_pending_comm_ops(0, True)
    release.set()
    time.sleep(0.1)
    assert "watchdog" in be.watchdog_error()
    be.shutdown()


def test_ready_spans_are_recorded():
    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    be.set_record_spans(True)
    ts = make(["p", "q"])
    b = C.Bucket("b", ts)
    be.register_ordered_buckets([b])
    be.mark_communication_ready(ts[1], 0)
    be.mark_communication_ready(ts[0], 0)
    be.wait_pending_comm_ops(0, True)
    spans = be.pop_ready_spans()
    assert [s[0] for s in spans] == ["q", "p"]
    be.shutdown()


def test_shift_one_peer_formula():
    # reference: decentralized_full_precision_synchronous.rs:81-85 (and tests/torch_api/test_decentralized.py:251-259)
    C = native()
    for n in (2, 4, 8):
        for step in range(6):
            peers = [C.PeerAverageOp.shift_one_peer(r, n, step) for r in range(n)]
            for r, p in enumerate(peers):
                assert peers[p] == r and p != r  # a perfect matching


def test_bucket_profile_measures_op_time_and_queueing():
    """Backend::set_profile on the CPU backend: host time of each bucket's op list, launch counts, table reset on re-registration."""
    import time

    C = native()
    backend = C.Backend(8, -1, 0, 60.0)
    t1 = C.Tensor("a", 1000, 16, 0, -1)
    t2 = C.Tensor("b", 2000, 32, 0, -1)
    b1, b2 = C.Bucket("slow", [t1]), C.Bucket("fast", [t2])
    b1.append_python_op(lambda name: time.sleep(0.03))
    b2.append_python_op(lambda name: None)
    backend.register_ordered_buckets([b1, b2])
    backend.set_profile(True)
    for _ in range(3):
        backend.mark_communication_ready(t1, 0)
        backend.mark_communication_ready(t2, 0)
        backend.wait_pending_comm_ops(0, True)
    stats = {s["name"]: s for s in backend.bucket_stats()}
    assert stats["slow"]["count"] == 3 and stats["fast"]["count"] == 3
    assert stats["slow"]["total_ms"] >= 80 and stats["slow"]["max_ms"] >= 25 and stats["fast"]["total_ms"] < 30
    assert stats["slow"]["bytes"] == 64 and stats["fast"]["bytes"] == 128 and "python" in stats["slow"]["ops"]
    assert backend.bucket_stats(True)[0]["count"] == 3 and backend.bucket_stats()[0]["count"] == 0   # reset
    backend.set_profile(False)
    backend.mark_communication_ready(t1, 0)
    backend.mark_communication_ready(t2, 0)
    backend.wait_pending_comm_ops(0, True)
    assert all(s["count"] == 0 for s in backend.bucket_stats())
    backend.shutdown()


def test_log_level_is_settable():
    C = native()
    before = C.log_level()
    C.set_log_level("debug")
    assert C.log_level() == "DEBUG"
    C.set_log_level("nonsense")
    assert C.log_level() == "WARN"
    C.set_log_level(before)


def test_inline_issue_runs_native_ops_on_the_marking_thread_and_keeps_the_order():
    """``Backend.set_inline(True)``: a bucket made of native ops only is issued by the thread that marks its last tensor (what
    a CUDA-graph capture needs: no second thread touches the streams); buckets with a python op still go through the worker,
    and a native bucket behind one of those follows it into the queue so the registered order is never violated."""
    import numpy as np

    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    src = [np.full(16, float(i + 1), dtype=np.float32) for i in range(3)]
    dst = [np.zeros(16, dtype=np.float32) for _ in range(3)]
    ts = [C.Tensor(f"t{i}", src[i].ctypes.data, 16, 0, -1) for i in range(3)]
    bs = [C.Bucket(f"b{i}", [ts[i]]) for i in range(3)]
    order, gate = [], threading.Event()
    bs[0].append_op(C.CopyOp(dst[0].ctypes.data, src[0].ctypes.data, 64))

    def slow_python_op(name):
        gate.wait(5)
        order.append(("python", bool(dst[2][0] == 3.0)))      # bucket 2 must not have been issued before this op finished

    bs[1].append_python_op(slow_python_op)
    bs[2].append_op(C.CopyOp(dst[2].ctypes.data, src[2].ctypes.data, 64))
    be.register_ordered_buckets(bs)
    assert not be.inline_mode()
    be.set_inline(True)
    be.mark_communication_ready(ts[0], 0)
    assert dst[0][0] == 1.0 and be.inline_total() == 1       # done before mark() returned: issued on this thread
    be.mark_communication_ready(ts[1], 0)                      # python op → worker thread (blocked on the gate)
    be.mark_communication_ready(ts[2], 0)                      # native, but the worker is busy → queued behind bucket 1
    assert dst[2][0] == 0.0 and be.inline_total() == 1
    gate.set()
    assert be.wait_pending_comm_ops(0, True) == 3
    assert order == [("python", False)] and dst[2][0] == 3.0
    # next iteration with the switch off: everything through the worker again
    be.set_inline(False)
    dst[0][:] = 0
    for t in ts:
        be.mark_communication_ready(t, 0)
    assert be.wait_pending_comm_ops(0, True) == 3 and dst[0][0] == 1.0 and be.inline_total() == 1
    be.shutdown()


def test_inline_issue_shares_the_profile_path_with_the_worker():
    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    t = C.Tensor("t", 0x1000, 4, 0, -1)
    b = C.Bucket("b", [t])

    b.append_op(C.CopyOp(0x10, 0x20, 0))      # zero-byte copy: a valid native op
    be.register_ordered_buckets([b])
    be.set_inline(True)
    be.set_profile(True)                      # the profile path is shared with the worker
    for _ in range(3):
        be.mark_communication_ready(t, 0)
        assert be.wait_pending_comm_ops(0, True) == 1
    stats = be.bucket_stats(False)
    assert stats[0]["count"] == 3 and be.inline_total() == 3
    be.shutdown()


def test_graph_capturable_only_with_native_step_invariant_ops():
    C = native()
    be = C.Backend(10, -1, 0, 30.0)
    t0, t1 = C.Tensor("a", 0x1000, 4, 0, -1), C.Tensor("b", 0x2000, 4, 0, -1)
    b0, b1 = C.Bucket("b0", [t0]), C.Bucket("b1", [t1])
    b0.append_op(C.CopyOp(0x10, 0x20, 0))
    be.register_ordered_buckets([b0, b1])
    assert be.graph_capturable()                      # native op + an empty program
    b1.append_python_op(lambda name: None)            # python ops run on the worker thread: not capturable
    assert not be.graph_capturable()
    be.shutdown()


def _timeline_worker(rank, world):
    """comm_timeline on the CPU backend (host timebase): every bucket execution is a sample, backward-end marks line up with the steps,
    the summary attributes the time communication ran on after backward, and the Chrome trace file carries all of it."""
    import json
    import os
    import tempfile
    import time

    import torch
    import torch.nn.functional as F

    import bagua_b200 as bagua
    from bagua_b200.parallel.algorithms import gradient_allreduce
    from bagua_b200.utils import export_chrome_trace

    bagua.init_process_group()
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    os.environ["BAGUA_DEFAULT_BUCKET_SIZE"] = "2048"     # several buckets
    net = net.with_bagua([opt], gradient_allreduce.GradientAllReduceAlgorithm())
    eng = net.bagua_ddp
    nb = len(net.bagua_buckets)
    for _ in range(2):                                      # before the switch: must not show up
        opt.zero_grad()
        F.mse_loss(net(torch.randn(8, 16)), torch.zeros(8, 4)).backward()
        opt.step()
    eng.comm_timeline(True)
    steps = 4
    for _ in range(steps):
        opt.zero_grad()
        F.mse_loss(net(torch.randn(8, 16)), torch.zeros(8, 4)).backward()
        opt.step()
    time.sleep(0.05)
    path = os.path.join(tempfile.mkdtemp(), f"trace{rank}.json")
    tl = export_chrome_trace(net, path)
    eng.comm_timeline(False)
    again = eng.comm_timeline_collect()
    with open(path) as f:
        trace = json.load(f)
    return nb, steps, tl, again, trace


def test_comm_timeline_and_chrome_trace_on_the_cpu_backend():
    from tests.mp_utils import run_distributed

    for nb, steps, tl, again, trace in run_distributed(_timeline_worker, world=2):
        assert nb >= 2
        assert len(tl["buckets"]) == nb * steps and len(tl["backward_end"]) == steps and len(tl["steps"]) == steps
        assert all(b["start_ms"] >= 0 and b["device_ms"] >= 0 and b["queue_ms"] >= 0 for b in tl["buckets"])
        per_iter = {}
        for b in tl["buckets"]:
            per_iter.setdefault(b["iteration"], []).append(b["bucket"])
        assert all(sorted(v) == sorted(per_iter[min(per_iter)]) and len(v) == nb for v in per_iter.values())
        marks = [m["ms"] for m in tl["backward_end"]]
        assert marks == sorted(marks) and [m["step"] for m in tl["backward_end"]] == list(range(3, 3 + steps))
        for s in tl["steps"]:
            assert s["buckets"] == nb and s["exposed_ms"] >= 0 and s["comm_end_ms"] >= 0 and abs(s["exposed_ms"] - max(0.0, s["comm_end_ms"] - s["backward_end_ms"])) < 1e-9
            assert s["comm_busy_ms"] >= 0 and s["last_bucket"] in per_iter[min(per_iter)]
        assert tl["ready"] and all(r["ms"] >= 0 for r in tl["ready"])
        assert len(tl["step_begin"]) == steps and again["buckets"] == [] and again["backward_end"] == [] and again["step_begin"] == []          # collected once, and nothing is recorded after the switch-off
        ev = trace["traceEvents"]
        assert sum(e["ph"] == "X" and e["tid"] == 1 for e in ev) == nb * steps and sum(e["ph"] == "i" and e["tid"] == 2 for e in ev) == 2 * steps
        assert any(e["ph"] == "M" and e["args"]["name"] == "comm stream" for e in ev) and len(trace["otherData"]["steps"]) == steps
        slices = [e for e in ev if e["ph"] == "X" and e["tid"] == 1]
        assert all("ops" in e["args"] and e["args"]["bytes"] > 0 for e in slices)


def test_summarize_timeline_attributes_exposed_time_to_the_right_step():
    from bagua_b200.parallel.bagua_distributed import summarize_timeline

    buckets = [  # two steps, two buckets each; step 1 hides its communication, step 2 leaves 3 ms exposed in bucket "1"
        {"bucket": "0", "iteration": 7, "start_ms": 1.0, "device_ms": 2.0}, {"bucket": "1", "iteration": 7, "start_ms": 4.0, "device_ms": 1.0},
        {"bucket": "0", "iteration": 8, "start_ms": 21.0, "device_ms": 2.0}, {"bucket": "1", "iteration": 8, "start_ms": 29.0, "device_ms": 4.0},
    ]
    begins = [{"step": 11, "ms": 20.0}, {"step": 10, "ms": 0.0}, {"step": 12, "ms": 40.0}]     # step 12: a forward without backward
    marks = [{"step": 11, "ms": 30.0}, {"step": 10, "ms": 10.0}]
    out = summarize_timeline(buckets, begins, marks)
    assert [(s["step"], s["exposed_ms"], s["last_bucket"], s["comm_busy_ms"]) for s in out] == [(10, 0.0, "1", 3.0), (11, 3.0, "1", 6.0)]
    assert summarize_timeline(buckets, begins, []) == [] and summarize_timeline(buckets, [], marks) == []


def test_timeline_gpu_branches_with_host_doubles(monkeypatch):
    """The CUDA-side bookkeeping of the timeline (events recorded in the hooks, resolved against the scheduler's reference event, events
    still in flight kept for the next collection) driven with stand-ins for ``torch.cuda.Event`` and the native backend."""
    import types

    import torch

    from bagua_b200.parallel.bagua_distributed import BaguaDistributedDataParallel as Engine

    class FakeEvent:
        made = []

        def __init__(self, enable_timing=False):
            assert enable_timing
            self.cuda_event, self.done, self.stream = 1000 + len(FakeEvent.made), True, None
            FakeEvent.made.append(self)

        def record(self, stream):
            self.stream = stream

        def query(self):
            return self.done

    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: "compute-stream")

    class FakeBackend:
        def __init__(self):
            self.calls = []

        def pop_bucket_timeline(self):
            return [{"bucket": "0", "iteration": 0, "issue_ns": 5, "start_ms": 2.0, "device_ms": 1.0, "queue_ms": 0.1},
                    {"bucket": "1", "iteration": 0, "issue_ns": 6, "start_ms": 4.0, "device_ms": 3.0, "queue_ms": 0.1}]

        def timeline_ref_ns(self):
            return 100

        def timeline_ms_of_event(self, ptr):
            self.calls.append(ptr)
            return {1000: 1.0, 1001: 5.0}.get(ptr, -1.0)     # begin at 1 ms, backward end at 5 ms; anything else unresolved

        def pop_ready_spans(self):
            return [("w", 50, 0), ("w", 2_000_100, 0)]       # the first one predates the reference

    eng = types.SimpleNamespace(_bagua_backend=FakeBackend(), _timeline_marks=[], _on_cuda=True, bagua_train_step_counter=7, _bagua_autotune_client=None)
    Engine._timeline_mark(eng, "begin")
    Engine._timeline_mark(eng, "backward_end")
    Engine._timeline_mark(eng, "begin")                      # next step has begun, its event is still in flight
    FakeEvent.made[2].done = False
    assert all(e.stream == "compute-stream" for e in FakeEvent.made) and [m[1] for m in eng._timeline_marks] == ["begin", "backward_end", "begin"]
    tl = Engine.comm_timeline_collect(eng)
    assert tl["step_begin"] == [{"step": 7, "ms": 1.0}] and tl["backward_end"] == [{"step": 7, "ms": 5.0}]
    assert tl["ready"] == [{"tensor": "w", "iteration": 0, "ms": 2.0}]
    assert len(eng._timeline_marks) == 1 and eng._timeline_marks[0][2] is FakeEvent.made[2]          # kept for the next collection
    (step,) = tl["steps"]
    assert step["step"] == 7 and step["exposed_ms"] == 2.0 and step["last_bucket"] == "1" and step["comm_busy_ms"] == 4.0 and step["buckets"] == 2
    assert eng._bagua_backend.calls == [1000, 1001]
