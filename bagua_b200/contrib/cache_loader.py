"""Read-through / write-behind sample cache shared by all data-loading workers.

Capability parity with the reference's ``CacheLoader`` (bagua/torch_api/contrib/cache_loader.py: ``CacheLoader(backend,
dataset_name, writer_buffer_size, **store_kwargs).get(key, load_fn)``); the implementation is organised differently: a
``_WriteBehind`` buffer owns the unflushed entries (flushed by count, on demand, and at teardown), lookups consult it
before the store, and the loader keeps hit / miss / error counters so a training script can see whether the cache works.
"""
from __future__ import annotations

import logging
import pickle
import weakref
from typing import Any, Callable, Dict, Optional

__all__ = ["CacheLoader", "BatchFetcher"]

logger = logging.getLogger(__name__)
_MISSING = object()


def serialize(input: Any) -> bytes:  # noqa: A002 - parameter name of the reference (cache_loader.py:9-14)
    """Wire format of cached values (pickle, highest protocol)."""
    return pickle.dumps(input, protocol=pickle.HIGHEST_PROTOCOL)


def deserialize(input: bytes) -> Any:  # noqa: A002
    return pickle.loads(input)


def _make_store(backend: str, options: dict):
    """The key-value store behind a loader: redis (cluster of per-node servers), the built-in TCP store, or process memory."""
    if backend == "memory":
        from .utils.store import MemoryStore

        return MemoryStore(), None
    if backend == "tcp":
        from .utils.store import TCPKVStore, start_tcp_kv_server

        if "port" in options:  # attach to a server somebody else started
            return TCPKVStore(options.get("host", "127.0.0.1"), options["port"]), None
        server, port = start_tcp_kv_server(0, options.get("capacity_per_node", 0))
        return TCPKVStore("127.0.0.1", port), server
    if backend == "redis":
        from .utils.redis_store import RedisStore

        return RedisStore(**options), None
    raise ValueError('Invalid backend, only support "redis", "tcp" and "memory"')


class _WriteBehind:
    """Entries waiting to be written: one ``mset`` per ``limit`` additions (or per explicit ``drain``). A failed ``mset``
    keeps the entries, so nothing computed by ``load_fn`` is lost while the store is briefly unreachable."""

    def __init__(self, store, limit: int):
        self.store = store
        self.limit = max(int(limit), 1)
        self.entries: Dict[str, bytes] = {}
        self.failed_flushes = 0

    def peek(self, key: str):
        blob = self.entries.get(key)
        return _MISSING if blob is None else deserialize(blob)

    def add(self, key: str, value: Any):
        self.entries[key] = serialize(value)
        if len(self.entries) >= self.limit:
            self.drain()

    def drain(self) -> bool:
        if not self.entries:
            return True
        try:
            self.store.mset(self.entries)
        except Exception as e:  # noqa: BLE001 - the cache is an optimisation: training continues without it
            self.failed_flushes += 1
            logger.debug("cache flush failed (%d entries kept): %s", len(self.entries), e)
            return False
        self.entries = {}
        return True


class BatchFetcher:
    """The reference's buffered store accessor (cache_loader.py:97-140: ``BatchFetcher(store, read_buffer_size, writer_buffer_size)``
    with ``read`` / ``write`` / ``write_post_read`` / ``flush_write_map``) on top of :class:`_WriteBehind`.  Values are python objects;
    a key that is still in the write buffer is served from there.  ``read_buffer_size`` is accepted for signature compatibility (reads
    go to the store one key at a time here: the stores answer a ``get`` in one round trip)."""

    def __init__(self, store, read_buffer_size: int = 1, writer_buffer_size: int = 1):
        self.store = store
        self.read_buffer_size = max(int(read_buffer_size), 1)
        self._pending = _WriteBehind(store, writer_buffer_size)

    @property
    def write_map(self) -> Dict[str, bytes]:
        return self._pending.entries

    def read(self, key: str):
        """The cached object, or ``None``."""
        value = self._pending.peek(key)
        if value is not _MISSING:
            return value
        blob = self.store.get(key)
        return None if blob is None else deserialize(blob)

    def write(self, key: str, value: Any):
        self._pending.add(key, value)

    def write_post_read(self):
        """Flush when the buffer has reached its size (``write`` already does; kept for call-site compatibility)."""
        if len(self._pending.entries) >= self._pending.limit:
            self._pending.drain()

    def flush_write_map(self):
        self._pending.drain()


class CacheLoader:
    r"""Caches the results of an expensive ``load_fn(key)`` (decode, tokenise, augment-once …) in a key-value store.

    Args:
        backend: ``"redis"`` (needs the ``redis`` package and server), ``"tcp"`` (built-in store) or ``"memory"``.
        dataset_name: key prefix, so several datasets can share one store.
        writer_buffer_size: number of new entries collected before they are written with one ``mset``.
        kwargs: forwarded to the store (``hosts``, ``cluster_mode``, ``capacity_per_node`` for redis; ``host``/``port`` for tcp).

    Example::

        >>> loader = CacheLoader(backend="tcp", dataset_name="imagenet-train", writer_buffer_size=32)
        >>> sample = loader.get(str(index), lambda k: decode(files[int(k)]))
    """

    def __init__(self, backend: str = "redis", dataset_name: str = "", writer_buffer_size: int = 1, **kwargs):
        self.backend = backend
        self.dataset_name = dataset_name
        self.store, self._server = _make_store(backend, kwargs)
        self._pending = _WriteBehind(self.store, writer_buffer_size)
        self.hits = self.misses = self.store_errors = 0
        self._finalizer = weakref.finalize(self, _WriteBehind.drain, self._pending)

    def _key(self, key) -> str:
        return f"{self.dataset_name}_{key}"

    def _lookup(self, full_key: str):
        value = self._pending.peek(full_key)
        if value is not _MISSING:
            return value
        try:
            blob = self.store.get(full_key)
        except Exception:  # noqa: BLE001
            self.store_errors += 1
            return _MISSING
        return _MISSING if blob is None else deserialize(blob)

    def get(self, key: str, load_fn: Callable[[str], Any]):
        """The value cached under ``key``; on a miss it is computed with ``load_fn(key)`` and queued for writing."""
        full_key = self._key(key)
        value = self._lookup(full_key)
        if value is not _MISSING:
            self.hits += 1
            return value
        self.misses += 1
        value = load_fn(key)
        self._pending.add(full_key, value)
        return value

    def flush(self) -> bool:
        """Write every queued entry now (also happens when the loader is garbage-collected)."""
        return self._pending.drain()

    def num_keys(self) -> int:
        """Entries visible in the store (queued ones are flushed first)."""
        self.flush()
        return self.store.num_keys()

    def stats(self) -> Dict[str, int]:
        return {"hits": self.hits, "misses": self.misses, "store_errors": self.store_errors, "queued": len(self._pending.entries),
                "failed_flushes": self._pending.failed_flushes}

    def __enter__(self) -> "CacheLoader":
        return self

    def __exit__(self, *exc) -> Optional[bool]:
        self.flush()
        return None
