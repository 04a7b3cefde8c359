"""Redis-backed cluster store (reference: bagua/torch_api/contrib/utils/redis_store.py:1-223).

Needs the ``redis`` python package and a ``redis-server`` binary; neither is a hard dependency.  ``RedisStore`` either
connects to the given ``hosts`` or bootstraps one ``redis-server`` per node and shares the endpoints through the c10d store."""
from __future__ import annotations

import logging
import subprocess
import time
from typing import Dict, List, Optional, Union

import torch.distributed as dist

from ... import env
from .store import ClusterStore, Store

__all__ = ["RedisStore"]

logger = logging.getLogger(__name__)
_host_ip = None
_server_proc = None
_BOOTSTRAP_PORT_KEY = "bagua_redis_port_node{}"


def _redis():
    try:
        import redis  # type: ignore

        return redis
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("RedisStore needs the `redis` package (and a redis-server binary); use backend 'memory' or 'tcp' instead") from e


class _RedisStore(Store):
    def __init__(self, host, port):
        self.host, self.port = host, port
        self.client = _redis().Redis(host=host, port=port)
        for _ in range(3):
            try:
                if self.client.ping():
                    break
            except Exception:  # noqa: BLE001
                time.sleep(1)

    def set(self, key, value):
        self.client.set(key, value)

    def get(self, key):
        return self.client.get(key)

    def num_keys(self) -> int:
        return self.client.dbsize()

    def clear(self):
        self.client.flushdb()

    def mset(self, dictionary: Dict[str, Union[str, bytes]]):
        self.client.mset(dictionary)

    def mget(self, keys: List[str]):
        return self.client.mget(keys)

    def status(self) -> bool:
        try:
            return bool(self.client.ping())
        except Exception:  # noqa: BLE001
            return False

    def shutdown(self):
        try:
            self.client.shutdown(nosave=True)
        except Exception:  # noqa: BLE001
            pass


def start_redis_server_cli(port: int, capacity: int, *args):
    cmd = ["redis-server", "--daemonize", "no", "--port", str(port), "--maxmemory", str(capacity), "--maxmemory-policy", "allkeys-random",
           "--appendonly", "no", "--save", "", "--protected-mode", "no", *args]
    return subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def bootstrap_redis_server(capacity_per_node: int) -> int:
    global _server_proc
    port = env.find_free_network_port()
    _server_proc = start_redis_server_cli(port, capacity_per_node)
    return port


def shutdown_redis_server():
    global _server_proc
    if _server_proc is not None:
        _server_proc.terminate()
        _server_proc = None


class RedisStore(ClusterStore):
    """
    Args:
        hosts: ``[{"host": ..., "port": ...}, ...]`` of existing redis servers, or ``None`` to bootstrap one per node.
        cluster_mode: shard over all nodes' servers (``True``) or use only the local node's server.
        capacity_per_node: memory cap of each bootstrapped server in bytes.
    """

    def __init__(self, hosts: Optional[List[Dict[str, str]]] = None, cluster_mode: bool = True, capacity_per_node: int = 107374182400):
        if hosts is None:
            local_rank, node_rank = env.get_local_rank(), env.get_node_rank()
            c10d_store = dist.distributed_c10d._get_default_store() if dist.is_initialized() else None
            if local_rank == 0:
                port = bootstrap_redis_server(capacity_per_node)
                if c10d_store is not None:
                    c10d_store.set(_BOOTSTRAP_PORT_KEY.format(node_rank), f"{env.get_master_addr() if node_rank == 0 else _get_host_ip()}:{port}")
                self._bootstrapped = True
            endpoints = []
            nnodes = max(1, env.get_world_size() // max(1, env.get_local_size()))
            nodes = range(nnodes) if cluster_mode else [node_rank]
            for n in nodes:
                host, port = c10d_store.get(_BOOTSTRAP_PORT_KEY.format(n)).decode().split(":") if c10d_store is not None else ("127.0.0.1", str(port))
                endpoints.append({"host": host, "port": int(port)})
            hosts = endpoints
        super().__init__([_RedisStore(h["host"], int(h["port"])) for h in hosts])


def _get_host_ip() -> str:
    import socket

    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    try:
        s.connect(("10.255.255.255", 1))
        return s.getsockname()[0]
    except Exception:  # noqa: BLE001
        return "127.0.0.1"
    finally:
        s.close()


def get_host_ip() -> str:
    """This host's outward-facing address (cached) — reference redis_store.py:216-223."""
    global _host_ip
    if _host_ip is None:
        _host_ip = _get_host_ip()
    return _host_ip


def create_redis_client(host: str, port: int):
    """A ``redis.Redis`` client; a server on this host is reached through loopback (reference redis_store.py:194-196)."""
    redis = _redis()
    return redis.Redis(port=port) if host in (get_host_ip(), "127.0.0.1", "localhost") else redis.Redis(host=host, port=port)
