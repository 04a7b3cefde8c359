"""Environment / configuration accessors (capability of the reference's bagua/torch_api/env.py).

Three tiers like the reference: launcher CLI flags → ``BAGUA_*`` / torch env vars → these accessors.  Every knob is one row
of ``SETTINGS`` (variable, type, default, meaning): the accessors below are thin named views of that table, and
``describe()`` renders it for ``--help`` texts and docs/configuration.md.
"""
from __future__ import annotations

import os
import socket
from contextlib import closing
from typing import Any, Callable, Dict, NamedTuple


class _Setting(NamedTuple):
    var: str
    cast: Callable[[str], Any]
    default: Any
    help: str


def _flag(v: str) -> bool:
    return v.strip().lower() not in ("", "0", "false", "no", "off")


SETTINGS: Dict[str, _Setting] = {s.var: s for s in (
    _Setting("WORLD_SIZE", int, 1, "number of processes in the job"),
    _Setting("RANK", int, 0, "global rank of this process"),
    _Setting("LOCAL_RANK", int, 0, "rank within the node"),
    _Setting("LOCAL_WORLD_SIZE", int, 1, "processes on this node"),
    _Setting("MASTER_ADDR", str, "127.0.0.1", "rendezvous host"),
    _Setting("MASTER_PORT", int, 29500, "rendezvous port"),
    _Setting("BAGUA_DEFAULT_BUCKET_SIZE", int, 10 * 1024 ** 2, "bucket size in bytes until autotune says otherwise"),
    _Setting("BAGUA_SERVICE_PORT", int, -1, "autotune service port (-1: pick a free one)"),
    _Setting("BAGUA_AUTOTUNE", int, 0, "autotune level (0 = off)"),
    _Setting("BAGUA_AUTOTUNE_MAX_SAMPLES", int, 60, "hyper-parameter samples before the tuner freezes the best one"),
    _Setting("BAGUA_AUTOTUNE_SAMPLING_CONFIDENCE_TIME_S", float, 5.0, "seconds a sample must run before its speed counts"),
    _Setting("BAGUA_AUTOTUNE_WARMUP_TIME_S", float, 30.0, "seconds before the first sample"),
    _Setting("BAGUA_IS_OUTPUT_AUTOTUNE_LOG", _flag, False, "write the tuner's CSV log"),
    _Setting("BAGUA_AUTOTUNE_SERVER_WAIT_TIME", int, 300, "seconds to wait for the autotune service to come up"),
    _Setting("BAGUA_REPORT_METRICS", _flag, False, "report training metrics"),
    _Setting("BAGUA_COMM_TIMEOUT_S", float, 300.0, "scheduler watchdog: limit for one bucket's communication"),
    _Setting("BAGUA_PEER_TIMEOUT_S", float, 300.0, "bound on in-kernel cross-GPU spins (same as the watchdog: a peer kernel that gives up is fatal)"),
    _Setting("BAGUA_ALLREDUCE_VARIANT", str.lower, "auto", "auto | one_shot | two_shot | multimem | nccl"),
)}


def _read(var: str):
    s = SETTINGS[var]
    raw = os.environ.get(var)
    return s.default if raw is None else s.cast(raw)


def describe() -> str:
    """One line per setting: ``VAR (default …): meaning``."""
    return "\n".join(f"{s.var} (default {s.default}): {s.help}" for s in SETTINGS.values())


# ---- launcher environment -------------------------------------------------------------------------------------------------
def get_world_size() -> int:
    return _read("WORLD_SIZE")


def get_rank() -> int:
    return _read("RANK")


def get_local_rank() -> int:
    return _read("LOCAL_RANK")


def get_local_size() -> int:
    return _read("LOCAL_WORLD_SIZE")


def get_master_addr() -> str:
    return _read("MASTER_ADDR")


def get_master_port() -> int:
    return _read("MASTER_PORT")


_ELASTIC_MARKERS = ("GROUP_RANK", "ROLE_RANK", "ROLE_WORLD_SIZE", "TORCHELASTIC_RESTART_COUNT")


def _is_elastic_launched() -> bool:
    """torch elastic exports these for every worker it starts."""
    return all(m in os.environ for m in _ELASTIC_MARKERS)


def get_node_rank() -> int:
    """Rank of the node: ``GROUP_RANK`` under the elastic launcher, else ``NODE_RANK``, else derived from rank / local size."""
    if _is_elastic_launched():
        return int(os.environ["GROUP_RANK"])
    explicit = os.environ.get("NODE_RANK")
    return int(explicit) if explicit is not None else get_rank() // max(get_local_size(), 1)


# ---- engine / autotune knobs -------------------------------------------------------------------------------------------------
def get_default_bucket_size() -> int:
    return _read("BAGUA_DEFAULT_BUCKET_SIZE")


def get_bagua_service_port() -> int:
    return _read("BAGUA_SERVICE_PORT")


def get_autotune_level() -> int:
    return _read("BAGUA_AUTOTUNE")


def get_autotune_max_samples() -> int:
    return _read("BAGUA_AUTOTUNE_MAX_SAMPLES")


def get_autotune_sampling_confidence_time_s() -> float:
    return _read("BAGUA_AUTOTUNE_SAMPLING_CONFIDENCE_TIME_S")


def get_autotune_warmup_time_s() -> float:
    return _read("BAGUA_AUTOTUNE_WARMUP_TIME_S")


def get_is_output_autotune_log() -> bool:
    return _read("BAGUA_IS_OUTPUT_AUTOTUNE_LOG")


def get_autotune_server_wait_time() -> int:
    return _read("BAGUA_AUTOTUNE_SERVER_WAIT_TIME")


def is_report_metrics_switch_on() -> bool:
    return _read("BAGUA_REPORT_METRICS")


def get_autotune_server_addr() -> str | None:
    return os.environ.get("AUTO_TUNE_SERVER_ADDR")


def get_comm_timeout_s() -> float:
    return _read("BAGUA_COMM_TIMEOUT_S")


def get_peer_kernel_timeout_s() -> float:
    return _read("BAGUA_PEER_TIMEOUT_S")


def get_allreduce_variant() -> str:
    return _read("BAGUA_ALLREDUCE_VARIANT")


def find_free_network_port() -> int:
    """A TCP port that is free right now on this host."""
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _document_accessors():
    """Every accessor that is a named view of one ``SETTINGS`` row gets that row as its docstring (``VAR``: meaning, default)."""
    import inspect
    import re

    for name, fn in list(globals().items()):
        if not inspect.isfunction(fn) or fn.__doc__ or fn.__module__ != __name__ or name.startswith("_"):
            continue
        try:
            m = re.search(r'_read\("([A-Z_]+)"\)', inspect.getsource(fn))
        except OSError:  # no source available (frozen / zipped install)
            m = None
        if m and m.group(1) in SETTINGS:
            st = SETTINGS[m.group(1)]
            fn.__doc__ = f"``{st.var}``: {st.help} (default {st.default!r})."


get_autotune_server_addr.__doc__ = "``AUTO_TUNE_SERVER_ADDR`` (``host:port`` of the autotune service; set by the launcher when autotune is on), or ``None``."
_document_accessors()
