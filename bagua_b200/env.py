"""Environment / configuration accessors (reference: bagua/torch_api/env.py:1-134).

Three tiers like the reference: launcher CLI flags → ``BAGUA_*`` / torch env vars → these accessors.
"""
from __future__ import annotations

import os
import socket
from contextlib import closing


def get_world_size() -> int:
    """Number of processes in the job (env ``WORLD_SIZE``, default 1)."""
    return int(os.environ.get("WORLD_SIZE", 1))


def get_rank() -> int:
    """Global rank of this process (env ``RANK``, default 0)."""
    return int(os.environ.get("RANK", 0))


def get_local_rank() -> int:
    """Rank within the node (env ``LOCAL_RANK``, default 0)."""
    return int(os.environ.get("LOCAL_RANK", 0))


def get_local_size() -> int:
    """Processes on this node (env ``LOCAL_WORLD_SIZE``, default 1)."""
    return int(os.environ.get("LOCAL_WORLD_SIZE", 1))


def _is_elastic_launched() -> bool:
    # torch elastic exports GROUP_RANK / TORCHELASTIC_RUN_ID (reference env.py:64-67)
    required = {"GROUP_RANK", "ROLE_RANK", "ROLE_WORLD_SIZE", "TORCHELASTIC_RESTART_COUNT"}
    return required.issubset(os.environ.keys())


def get_node_rank() -> int:
    """Rank of the node; comes from ``GROUP_RANK`` under the elastic launcher (reference env.py:51-67)."""
    if _is_elastic_launched():
        return int(os.environ.get("GROUP_RANK", 0))
    if "NODE_RANK" in os.environ:
        return int(os.environ["NODE_RANK"])
    local = max(get_local_size(), 1)
    return get_rank() // local


def get_master_addr() -> str:
    return os.environ.get("MASTER_ADDR", "127.0.0.1")


def get_master_port() -> int:
    return int(os.environ.get("MASTER_PORT", 29500))


def get_default_bucket_size() -> int:
    """Bucket size in bytes used until autotune says otherwise (10 MiB, reference env.py:70-76)."""
    return int(os.environ.get("BAGUA_DEFAULT_BUCKET_SIZE", 10 * 1024 ** 2))


def get_bagua_service_port() -> int:
    return int(os.environ.get("BAGUA_SERVICE_PORT", -1))


def get_autotune_level() -> int:
    return int(os.environ.get("BAGUA_AUTOTUNE", 0))


def get_autotune_max_samples() -> int:
    return int(os.environ.get("BAGUA_AUTOTUNE_MAX_SAMPLES", 60))


def get_autotune_sampling_confidence_time_s() -> float:
    return float(os.environ.get("BAGUA_AUTOTUNE_SAMPLING_CONFIDENCE_TIME_S", 5.0))


def get_autotune_warmup_time_s() -> float:
    return float(os.environ.get("BAGUA_AUTOTUNE_WARMUP_TIME_S", 30.0))


def get_is_output_autotune_log() -> bool:
    return bool(int(os.environ.get("BAGUA_IS_OUTPUT_AUTOTUNE_LOG", 0)))


def get_autotune_server_wait_time() -> int:
    return int(os.environ.get("BAGUA_AUTOTUNE_SERVER_WAIT_TIME", 300))


def is_report_metrics_switch_on() -> bool:
    return int(os.environ.get("BAGUA_REPORT_METRICS", 0)) == 1


def get_autotune_server_addr() -> str | None:
    return os.environ.get("AUTO_TUNE_SERVER_ADDR")


def get_comm_timeout_s() -> float:
    """Watchdog limit for one bucket's communication (reference: 300 s, lib.rs:259)."""
    return float(os.environ.get("BAGUA_COMM_TIMEOUT_S", 300.0))


def get_peer_kernel_timeout_s() -> float:
    """Bound on in-kernel cross-GPU spins; keeps a lost peer from hanging the GPU."""
    return float(os.environ.get("BAGUA_PEER_TIMEOUT_S", 60.0))


def get_allreduce_variant() -> str:
    """``auto`` | ``one_shot`` | ``two_shot`` | ``multimem`` | ``nccl`` (fallback / baseline)."""
    return os.environ.get("BAGUA_ALLREDUCE_VARIANT", "auto").lower()


def find_free_network_port() -> int:
    """A TCP port that is free right now on this host (reference env.py:125-134)."""
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.bind(("127.0.0.1", 0))
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        return s.getsockname()[1]
