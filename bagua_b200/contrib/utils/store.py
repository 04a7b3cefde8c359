"""Key-value stores backing :class:`~bagua_b200.contrib.CacheLoader`
(reference: bagua/torch_api/contrib/utils/store.py:1-145)."""
from __future__ import annotations

import pickle
import socket
import socketserver
import struct
import threading
from typing import Dict, List, Optional, Union

__all__ = ["Store", "ClusterStore", "MemoryStore", "TCPKVStore", "start_tcp_kv_server"]

Value = Union[str, bytes]


class Store:
    """Base class of key-value stores: ``set/get`` single entries, ``mset/mget`` batches."""

    def set(self, key: str, value: Value):
        raise NotImplementedError

    def get(self, key: str) -> Optional[Value]:
        raise NotImplementedError

    def num_keys(self) -> int:
        raise NotImplementedError

    def clear(self):
        raise NotImplementedError

    def mset(self, dictionary: Dict[str, Value]):
        for k, v in dictionary.items():
            self.set(k, v)

    def mget(self, keys: List[str]) -> List[Optional[Value]]:
        return [self.get(k) for k in keys]

    def status(self) -> bool:
        raise NotImplementedError

    def shutdown(self):
        raise NotImplementedError


class MemoryStore(Store):
    """In-process dictionary store (single-process jobs, tests)."""

    def __init__(self):
        self._d: Dict[str, Value] = {}
        self._lock = threading.Lock()

    def set(self, key, value):
        with self._lock:
            self._d[key] = value

    def get(self, key):
        with self._lock:
            return self._d.get(key)

    def num_keys(self):
        return len(self._d)

    def clear(self):
        with self._lock:
            self._d.clear()

    def status(self):
        return True

    def shutdown(self):
        self.clear()


class ClusterStore(Store):
    """Shards entries over several stores by the xxh64 hash of the key."""

    def __init__(self, stores: List[Store]):
        self.stores = stores
        self.num_stores = len(stores)
        try:
            import xxhash

            self.hash_fn = lambda b: xxhash.xxh64(b).intdigest()
        except ImportError:  # pragma: no cover
            import zlib

            self.hash_fn = lambda b: zlib.crc32(b)

    def _hash_key(self, key: str) -> int:
        return self.hash_fn(key.encode()) % self.num_stores

    def route(self, key: str) -> Store:
        return self.stores[self._hash_key(key)] if self.num_stores > 1 else self.stores[0]

    def set(self, key, value):
        self.route(key).set(key, value)

    def get(self, key):
        return self.route(key).get(key)

    def num_keys(self) -> int:
        return sum(s.num_keys() for s in self.stores)

    def clear(self):
        for s in self.stores:
            s.clear()

    def mset(self, dictionary):
        if self.num_stores == 1:
            return self.stores[0].mset(dictionary)
        table: Dict[int, Dict[str, Value]] = {}
        for k, v in dictionary.items():
            table.setdefault(self._hash_key(k), {})[k] = v
        for sid, m in table.items():
            self.stores[sid].mset(m)

    def mget(self, keys):
        if self.num_stores == 1:
            return self.stores[0].mget(keys)
        table: Dict[int, List[str]] = {}
        for k in keys:
            table.setdefault(self._hash_key(k), []).append(k)
        found: Dict[str, Optional[Value]] = {}
        for sid, ks in table.items():
            found.update(zip(ks, self.stores[sid].mget(ks)))
        return [found.get(k) for k in keys]

    def status(self) -> bool:
        return all(s.status() for s in self.stores)

    def shutdown(self):
        for s in self.stores:
            s.shutdown()


# ---- a dependency-free networked store (one server per node), used when redis is not installed ------------------
def _send_msg(sock, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("!Q", len(data)) + data)


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("kv store connection closed")
        buf.extend(chunk)
    return bytes(buf)


def _recv_msg(sock):
    (n,) = struct.unpack("!Q", _recv_exact(sock, 8))
    return pickle.loads(_recv_exact(sock, n))


class _KVHandler(socketserver.BaseRequestHandler):
    def handle(self):
        srv = self.server
        while True:
            try:
                op, arg = _recv_msg(self.request)
            except (ConnectionError, OSError, struct.error):
                return
            with srv.lock:
                if op == "mset":
                    for k, v in arg.items():
                        if srv.capacity and srv.bytes_used + len(v) > srv.capacity and k not in srv.data:
                            continue  # full: behave like a cache, drop new entries
                        srv.bytes_used += len(v) - len(srv.data.get(k, b""))
                        srv.data[k] = v
                    out = True
                elif op == "mget":
                    out = [srv.data.get(k) for k in arg]
                elif op == "num_keys":
                    out = len(srv.data)
                elif op == "clear":
                    srv.data.clear()
                    srv.bytes_used = 0
                    out = True
                elif op == "ping":
                    out = True
                elif op == "shutdown":
                    out = True
                else:
                    out = None
            _send_msg(self.request, out)
            if op == "shutdown":
                threading.Thread(target=srv.shutdown, daemon=True).start()
                return


class _KVServer(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True

    def __init__(self, addr, capacity):
        super().__init__(addr, _KVHandler)
        self.data: Dict[str, bytes] = {}
        self.lock = threading.Lock()
        self.capacity = capacity
        self.bytes_used = 0


def start_tcp_kv_server(port: int = 0, capacity_bytes: int = 0, host: str = "0.0.0.0"):
    """Start a KV server thread in this process; returns ``(server, port)``."""
    srv = _KVServer((host, port), capacity_bytes)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    return srv, srv.server_address[1]


class TCPKVStore(Store):
    """Client of :func:`start_tcp_kv_server`."""

    def __init__(self, host: str, port: int):
        self.host, self.port = host, port
        self._local = threading.local()

    def _sock(self):
        s = getattr(self._local, "sock", None)
        if s is None:
            s = socket.create_connection((self.host, self.port), timeout=30)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            self._local.sock = s
        return s

    def _call(self, op, arg=None):
        try:
            s = self._sock()
            _send_msg(s, (op, arg))
            return _recv_msg(s)
        except (ConnectionError, OSError):
            self._local.sock = None
            raise

    @staticmethod
    def _b(v: Value) -> bytes:
        return v if isinstance(v, bytes) else str(v).encode()

    def set(self, key, value):
        self._call("mset", {key: self._b(value)})

    def get(self, key):
        return self._call("mget", [key])[0]

    def mset(self, dictionary):
        self._call("mset", {k: self._b(v) for k, v in dictionary.items()})

    def mget(self, keys):
        return self._call("mget", list(keys))

    def num_keys(self):
        return self._call("num_keys")

    def clear(self):
        self._call("clear")

    def status(self):
        try:
            return bool(self._call("ping"))
        except Exception:  # noqa: BLE001
            return False

    def shutdown(self):
        try:
            self._call("shutdown")
        except Exception:  # noqa: BLE001
            pass
