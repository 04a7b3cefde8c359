"""Python side of the NCCL network plugin (``libnccl-net-bagua.so``, sources in ``csrc/net``).

Counterpart of the reference's Bagua-Net switch (``--enable_bagua_net`` in bagua/distributed/launch.py:102-107 sets
``LD_LIBRARY_PATH`` so that NCCL finds ``libnccl-net.so``).  Here :func:`enable` exports ``NCCL_NET_PLUGIN=bagua`` plus the
library directory, and :class:`PluginHandle` drives the plugin's function table through ``ctypes`` — used by the CPU tests
and by ``python -m bagua_b200.net`` (a loopback throughput probe) without involving NCCL or a GPU.

Only inter-node traffic ever reaches a network plugin; inside one NVSwitch domain NCCL and the peer-memory kernels bypass
it, so enabling it on a single node is harmless and pointless.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from pathlib import Path
from typing import List, Optional, Tuple

PLUGIN_NAME = "bagua"
_LIB = Path(__file__).resolve().parent / "libnccl-net-bagua.so"
HANDLE_BYTES = 128


def plugin_path(build: bool = True) -> Path:
    """Path of ``libnccl-net-bagua.so`` (built on demand)."""
    if build and not _LIB.exists():
        from ._build import build_net_plugin

        build_net_plugin()
    return _LIB


def enable(env: Optional[dict] = None, nstreams: Optional[int] = None, min_chunksize: Optional[int] = None) -> dict:
    """Make NCCL load the plugin in processes that inherit ``env`` (default: ``os.environ``).  Must run before the first
    NCCL communicator is created."""
    env = os.environ if env is None else env
    lib = plugin_path()
    env["NCCL_NET_PLUGIN"] = PLUGIN_NAME
    parts = [str(lib.parent)] + [p for p in env.get("LD_LIBRARY_PATH", "").split(":") if p]
    env["LD_LIBRARY_PATH"] = ":".join(dict.fromkeys(parts))
    if nstreams is not None:
        env["BAGUA_NET_NSTREAMS"] = str(int(nstreams))
    if min_chunksize is not None:
        env["BAGUA_NET_MIN_CHUNKSIZE"] = str(int(min_chunksize))
    return env


class _Props(C.Structure):
    _fields_ = [("name", C.c_char_p), ("pciPath", C.c_char_p), ("guid", C.c_uint64), ("ptrSupport", C.c_int), ("speed", C.c_int),
                ("port", C.c_int), ("latency", C.c_float), ("maxComms", C.c_int), ("maxRecvs", C.c_int)]


_VP, _VPP, _IP = C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)


class _NetV6(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("init", C.CFUNCTYPE(C.c_int, _VP)),
        ("devices", C.CFUNCTYPE(C.c_int, _IP)),
        ("getProperties", C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(_Props))),
        ("listen", C.CFUNCTYPE(C.c_int, C.c_int, _VP, _VPP)),
        ("connect", C.CFUNCTYPE(C.c_int, C.c_int, _VP, _VPP)),
        ("accept", C.CFUNCTYPE(C.c_int, _VP, _VPP)),
        ("regMr", C.CFUNCTYPE(C.c_int, _VP, _VP, C.c_int, C.c_int, _VPP)),
        ("regMrDmaBuf", C.CFUNCTYPE(C.c_int, _VP, _VP, C.c_size_t, C.c_int, C.c_uint64, C.c_int, _VPP)),
        ("deregMr", C.CFUNCTYPE(C.c_int, _VP, _VP)),
        ("isend", C.CFUNCTYPE(C.c_int, _VP, _VP, C.c_int, C.c_int, _VP, _VPP)),
        ("irecv", C.CFUNCTYPE(C.c_int, _VP, C.c_int, _VPP, _IP, _IP, _VPP, _VPP)),
        ("iflush", C.CFUNCTYPE(C.c_int, _VP, C.c_int, _VPP, _IP, _VPP, _VPP)),
        ("test", C.CFUNCTYPE(C.c_int, _VP, _IP, _IP)),
        ("closeSend", C.CFUNCTYPE(C.c_int, _VP)),
        ("closeRecv", C.CFUNCTYPE(C.c_int, _VP)),
        ("closeListen", C.CFUNCTYPE(C.c_int, _VP)),
    ]


class PluginHandle:
    """The plugin's ``ncclNetPlugin_v6`` table, callable from Python exactly the way NCCL's proxy thread calls it."""

    def __init__(self):
        self.lib = C.CDLL(str(plugin_path()))
        self.net = _NetV6.in_dll(self.lib, "ncclNetPlugin_v6")
        rc = self.net.init(None)
        if rc != 0:
            raise RuntimeError(f"plugin init failed (ncclResult {rc}); NCCL_SOCKET_IFNAME={os.environ.get('NCCL_SOCKET_IFNAME')}")

    @property
    def name(self) -> str:
        return self.net.name.decode()

    def devices(self) -> List[dict]:
        n = C.c_int(0)
        assert self.net.devices(C.byref(n)) == 0
        out = []
        for d in range(n.value):
            p = _Props()
            assert self.net.getProperties(d, C.byref(p)) == 0
            out.append({"name": p.name.decode(), "speed_mbps": p.speed, "ptr_support": p.ptrSupport, "max_recvs": p.maxRecvs,
                        "pci_path": p.pciPath.decode() if p.pciPath else None})
        return out

    def listen(self, dev: int = 0) -> Tuple[C.Array, C.c_void_p]:
        handle = (C.c_char * HANDLE_BYTES)()
        comm = C.c_void_p()
        rc = self.net.listen(dev, C.cast(handle, _VP), C.byref(comm))
        if rc != 0:
            raise RuntimeError(f"listen failed: {rc}")
        return handle, comm

    def connect(self, handle, dev: int = 0) -> C.c_void_p:
        comm = C.c_void_p()
        rc = self.net.connect(dev, C.cast(handle, _VP), C.byref(comm))
        if rc != 0 or not comm.value:
            raise RuntimeError(f"connect failed: {rc}")
        return comm

    def accept(self, listen_comm, timeout_s: float = 20.0) -> C.c_void_p:
        comm = C.c_void_p()
        deadline = time.time() + timeout_s
        while time.time() < deadline:
            rc = self.net.accept(listen_comm, C.byref(comm))
            if rc != 0:
                raise RuntimeError(f"accept failed: {rc}")
            if comm.value:
                return comm
            time.sleep(0.001)
        raise TimeoutError("accept: peer streams did not arrive")

    def isend(self, comm, buf, nbytes: int) -> Optional[C.c_void_p]:
        req = C.c_void_p()
        rc = self.net.isend(comm, C.cast(buf, _VP), nbytes, 0, None, C.byref(req))
        if rc != 0:
            raise RuntimeError(f"isend failed: {rc}")
        return req if req.value else None

    def irecv(self, comm, buf, capacity: int) -> Optional[C.c_void_p]:
        req = C.c_void_p()
        data = (C.c_void_p * 1)(C.cast(buf, _VP))
        sizes = (C.c_int * 1)(capacity)
        tags = (C.c_int * 1)(0)
        mh = (C.c_void_p * 1)(None)
        rc = self.net.irecv(comm, 1, data, sizes, tags, mh, C.byref(req))
        if rc != 0:
            raise RuntimeError(f"irecv failed: {rc}")
        return req if req.value else None

    def test(self, req) -> Tuple[bool, int, int]:
        done, size = C.c_int(0), C.c_int(0)
        rc = self.net.test(req, C.byref(done), C.byref(size))
        return bool(done.value), size.value, rc

    def wait(self, req, timeout_s: float = 30.0) -> int:
        deadline = time.time() + timeout_s
        while time.time() < deadline:
            done, size, rc = self.test(req)
            if done:
                if rc != 0:
                    raise RuntimeError(f"request failed: ncclResult {rc}")
                return size
        raise TimeoutError("request did not complete")

    def close_send(self, comm):
        self.net.closeSend(comm)

    def close_recv(self, comm):
        self.net.closeRecv(comm)

    def close_listen(self, comm):
        self.net.closeListen(comm)

    def plan_chunks(self, size: int, nstreams: int, min_chunk: int, cursor: int = 0) -> List[Tuple[int, int, int]]:
        cap = 64
        out = (C.c_uint64 * (3 * cap))()
        self.lib.bagua_net_plan_chunks.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint, C.POINTER(C.c_uint64), C.c_int]
        n = self.lib.bagua_net_plan_chunks(size, nstreams, min_chunk, cursor, out, cap)
        return [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(n)]

    def stats(self) -> dict:
        buf = C.create_string_buffer(1024)
        self.lib.bagua_net_stats_json(buf, 1024)
        return json.loads(buf.value.decode())

    def trace_flush(self):
        """Write buffered isend/irecv spans to the trace file (``BAGUA_NET_TRACE_FILE`` / ``BAGUA_NET_JAEGER_ADDRESS``)."""
        self.lib.bagua_net_trace_flush()


def loopback_probe(nbytes: int = 256 << 20, iters: int = 8) -> dict:
    """Throughput of one connection over the first usable interface (both ends in this process)."""
    import numpy as np

    net = PluginHandle()
    handle, lc = net.listen(0)
    sc = net.connect(handle)
    rc = net.accept(lc)
    src = np.random.default_rng(0).integers(0, 255, nbytes, dtype=np.uint8)
    dst = np.zeros(nbytes, dtype=np.uint8)
    t0 = time.perf_counter()
    for _ in range(iters):
        r = net.irecv(rc, dst.ctypes.data, nbytes)
        s = net.isend(sc, src.ctypes.data, nbytes)
        net.wait(s)
        net.wait(r)
    dt = time.perf_counter() - t0
    ok = bool((src == dst).all())
    net.close_send(sc), net.close_recv(rc), net.close_listen(lc)
    return {"device": net.devices()[0], "bytes": nbytes, "iters": iters, "GB_per_s": nbytes * iters / dt / 1e9, "intact": ok, "stats": net.stats()}


if __name__ == "__main__":
    print(json.dumps(loopback_probe(), indent=1))
