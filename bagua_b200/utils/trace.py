"""Chrome / Perfetto trace of one module's communication (``chrome://tracing`` or https://ui.perfetto.dev open the file).

    model.bagua_ddp.comm_timeline(True)
    ... a few training steps ...; torch.cuda.synchronize()
    export_chrome_trace(model, "comm.trace.json")

Tracks: ``comm stream`` — one slice per execution of a bucket's op list (name = bucket, args = ops, bytes, host queueing delay);
``compute stream`` — an instant mark where backward ended in each step and, from there, a slice ``exposed communication`` up to the
end of the last bucket when communication was not fully hidden; ``host`` — the tensor-ready marks of the scheduler.  The reference
ships per-tensor spans to its autotune service only (bagua-opentelemetry); there is no per-bucket device timeline in it."""
from __future__ import annotations

import json
from typing import Optional

__all__ = ["export_chrome_trace", "timeline_to_trace_events"]


def _engine(obj):
    for attr in ("bagua_ddp", "inner"):
        if hasattr(obj, attr):
            return getattr(obj, attr)
    return obj


def timeline_to_trace_events(timeline: dict, pid: int = 0, programs: Optional[dict] = None) -> list:
    """Trace-event dictionaries for the output of ``BaguaDistributedDataParallel.comm_timeline_collect()``."""
    programs = programs or {}
    ev = [{"ph": "M", "pid": pid, "tid": tid, "name": "thread_name", "args": {"name": name}}
          for tid, name in ((1, "comm stream"), (2, "compute stream"), (3, "host (tensor ready)"))]
    for b in timeline["buckets"]:
        info = programs.get(b["bucket"], {})
        ev.append({"ph": "X", "pid": pid, "tid": 1, "name": f"bucket {b['bucket']}", "ts": b["start_ms"] * 1e3, "dur": max(b["device_ms"], 0.0) * 1e3,
                   "args": {"iteration": b["iteration"], "queue_ms": b["queue_ms"], **info}})
    for m in timeline.get("step_begin", []):
        ev.append({"ph": "i", "s": "t", "pid": pid, "tid": 2, "name": f"step {m['step']} begins", "ts": m["ms"] * 1e3})
    for m in timeline["backward_end"]:
        ev.append({"ph": "i", "s": "t", "pid": pid, "tid": 2, "name": f"backward end (step {m['step']})", "ts": m["ms"] * 1e3})
    for s in timeline["steps"]:
        if s["exposed_ms"] > 0:
            ev.append({"ph": "X", "pid": pid, "tid": 2, "name": "exposed communication", "ts": s["backward_end_ms"] * 1e3, "dur": s["exposed_ms"] * 1e3,
                       "args": {"step": s["step"], "last_bucket": s["last_bucket"], "comm_busy_ms": s["comm_busy_ms"]}})
    for r in timeline["ready"]:
        ev.append({"ph": "i", "s": "t", "pid": pid, "tid": 3, "name": r["tensor"], "ts": r["ms"] * 1e3, "args": {"iteration": r["iteration"]}})
    return ev


def export_chrome_trace(module_or_engine, path: str, rank: Optional[int] = None) -> dict:
    """Collect the timeline of ``module_or_engine`` (a module returned by ``with_bagua``, a DDP wrapper or the engine itself), write it as a
    trace-event JSON file and return the collected dictionary (its ``steps`` list carries the exposed milliseconds per step)."""
    from .. import env

    eng = _engine(module_or_engine)
    timeline = eng.comm_timeline_collect()
    programs = {r["bucket"]: {"ops": r["ops"], "bytes": r["bytes"], "variant": r["variant"]} for r in eng.comm_report()}
    pid = env.get_rank() if rank is None else rank
    with open(path, "w") as f:
        json.dump({"traceEvents": timeline_to_trace_events(timeline, pid, programs), "displayTimeUnit": "ms",
                   "otherData": {"module": eng.bagua_module_name, "rank": pid, "steps": timeline["steps"]}}, f)
    return timeline
