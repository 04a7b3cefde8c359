"""Read-through sample cache (reference: bagua/torch_api/contrib/cache_loader.py:1-140)."""
from __future__ import annotations

import pickle
from typing import Callable, Dict

__all__ = ["CacheLoader"]


def serialize(obj) -> bytes:
    return pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)


def deserialize(data):
    return pickle.loads(data)


class CacheLoader:
    r"""Caches the results of an expensive ``load_fn(key)`` in a key-value store shared by all data-loading workers.

    Args:
        backend: ``"redis"`` (needs redis), ``"tcp"`` (built-in one-server-per-process store) or ``"memory"``.
        dataset_name: key prefix, so several datasets can share one store.
        writer_buffer_size: number of writes batched into one ``mset``.
        kwargs: forwarded to the store constructor (e.g. ``hosts``, ``cluster_mode``, ``capacity_per_node`` for redis).
    """

    def __init__(self, backend: str = "redis", dataset_name: str = "", writer_buffer_size: int = 1, **kwargs):
        self.backend = backend
        self.dataset_name = dataset_name
        if backend == "redis":
            from .utils.redis_store import RedisStore

            self.store = RedisStore(**kwargs)
        elif backend == "memory":
            from .utils.store import MemoryStore

            self.store = MemoryStore()
        elif backend == "tcp":
            from .utils.store import TCPKVStore, start_tcp_kv_server

            if "port" in kwargs:
                self.store = TCPKVStore(kwargs.get("host", "127.0.0.1"), kwargs["port"])
            else:
                self._server, port = start_tcp_kv_server(0, kwargs.get("capacity_per_node", 0))
                self.store = TCPKVStore("127.0.0.1", port)
        else:
            raise ValueError('Invalid backend, only support "redis", "tcp" and "memory"')
        self.fetcher = BatchFetcher(self.store, 1, writer_buffer_size)

    def get(self, key: str, load_fn: Callable[[str], object]):
        """Value cached under ``key``; computed with ``load_fn(key)`` and stored on a miss."""
        cache_key = f"{self.dataset_name}_{key}"
        ret = self.fetcher.read(cache_key)
        if ret is None:
            ret = load_fn(key)
            self.fetcher.write(cache_key, ret)
        return ret

    def num_keys(self) -> int:
        return self.store.num_keys()


class BatchFetcher:
    def __init__(self, store, read_buffer_size: int, writer_buffer_size: int):
        self.store = store
        self.read_buffer_size = max(1, read_buffer_size)
        self.writer_buffer_size = max(1, writer_buffer_size)
        self.write_map: Dict[str, bytes] = {}
        self.write_cnt = 0
        self.read_cnt = 0

    def read(self, key):
        self.read_cnt += 1
        if key in self.write_map:  # not flushed yet
            return deserialize(self.write_map[key])
        try:
            ret = self.store.get(key)
        except Exception:  # noqa: BLE001
            ret = None
        else:
            if self.read_cnt % 1000 == 0 and self.write_map:
                self.flush_write_map()
        return deserialize(ret) if ret is not None else None

    def write(self, key, value):
        self.write_cnt += 1
        self.write_map[key] = serialize(value)
        if self.write_cnt % self.writer_buffer_size == 0:
            self.flush_write_map()

    def flush_write_map(self):
        try:
            self.store.mset(self.write_map)
        except Exception:  # noqa: BLE001
            pass
        else:
            self.write_map.clear()
