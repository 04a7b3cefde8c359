"""NCCL network plugin (csrc/net): the v6 function table is driven over loopback exactly as NCCL's proxy would.

The reference unit-tests only URL/sockaddr parsing and chunking of its plugin (rust/bagua-net/src/utils.rs:268-314);
this covers the whole data path — connection set-up, inline and multi-stream messages, ordering, back-pressure, shutdown.
"""
import ctypes as C
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def net():
    os.environ["NCCL_SOCKET_IFNAME"] = "lo"      # the sandbox may have no other interface
    os.environ["BAGUA_NET_NSTREAMS"] = "3"
    os.environ["BAGUA_NET_MIN_CHUNKSIZE"] = "65536"
    from bagua_b200.net import PluginHandle

    return PluginHandle()


def test_plan_chunks_is_a_partition(net):
    for size, streams, min_chunk, cursor in [(1, 4, 1 << 20, 0), (5 << 20, 4, 1 << 20, 3), (1 << 20, 2, 1 << 20, 1), (10_000_001, 8, 4096, 5),
                                             (3 << 20, 16, 1 << 20, 0)]:
        plan = net.plan_chunks(size, streams, min_chunk, cursor)
        assert 1 <= len(plan) <= streams and len(plan) <= max(1, size // min_chunk)
        off = 0
        for i, (o, b, s) in enumerate(plan):
            assert o == off and b > 0 and s == (cursor + i) % streams
            assert o % 16 == 0
            off += b
        assert off == size
        sizes = [b for _, b, _ in plan]
        assert max(sizes) - min(sizes) <= 16 * len(plan) + (size % 16)      # balanced shares
    assert net.plan_chunks(0, 4, 4096) == []


def test_devices_and_properties(net):
    devs = net.devices()
    assert net.name == "BaguaNet-B200" and len(devs) >= 1
    assert devs[0]["name"] == "lo" and devs[0]["ptr_support"] == 1 and devs[0]["max_recvs"] == 1 and devs[0]["speed_mbps"] > 0


def test_messages_of_every_size_class_arrive_intact_and_in_order(net):
    handle, lc = net.listen(0)
    sc = net.connect(handle)
    rc = net.accept(lc)
    rng = np.random.default_rng(1)
    before = net.stats()
    sizes = [0, 1, 100, 16 * 1024, 16 * 1024 + 1, 65536, 200_000, 1_000_003, 7 << 20]
    srcs = [rng.integers(0, 256, s, dtype=np.uint8) for s in sizes]
    dsts = [np.full(max(s, 1) + 64, 0xEE, dtype=np.uint8) for s in sizes]       # receive buffers are larger than the message
    recv_reqs, send_reqs = [], []
    for s, d in zip(sizes, dsts):                                               # all receives posted up front
        recv_reqs.append(net.irecv(rc, d.ctypes.data, d.nbytes))
    for s, a in zip(sizes, srcs):
        send_reqs.append(net.isend(sc, a.ctypes.data if s else None, s))
    assert all(r is not None for r in recv_reqs + send_reqs)
    for r, s in zip(send_reqs, sizes):
        assert net.wait(r) == s
    for r, s, a, d in zip(recv_reqs, sizes, srcs, dsts):
        assert net.wait(r) == s                                                 # the receiver learns the real size
        np.testing.assert_array_equal(d[:s], a)
        assert (d[s:] == 0xEE).all()                                            # nothing written past the message
    after = net.stats()
    assert after["bytes_sent"] - before["bytes_sent"] == sum(sizes) == after["bytes_received"] - before["bytes_received"]
    assert after["chunks"] > before["chunks"] and after["inline_msgs"] > before["inline_msgs"]
    # a message larger than the posted buffer is an error on the receiving side, not a buffer overrun
    small = np.zeros(1000, dtype=np.uint8)
    big = rng.integers(0, 256, 100_000, dtype=np.uint8)
    r = net.irecv(rc, small.ctypes.data, small.nbytes)
    s = net.isend(sc, big.ctypes.data, big.nbytes)
    with pytest.raises(RuntimeError):
        net.wait(r)
    net.close_send(sc)
    net.close_recv(rc)
    net.close_listen(lc)
    del s


def test_request_pool_back_pressure_and_two_connections(net):
    handle, lc = net.listen(0)
    sc1, sc2 = net.connect(handle), net.connect(handle)
    rc1, rc2 = net.accept(lc), net.accept(lc)
    # 64 request slots per connection: the 65th un-tested isend reports "try again" (request == NULL), never blocks
    payload = np.arange(256, dtype=np.uint8)
    reqs = []
    for _ in range(70):
        r = net.isend(sc1, payload.ctypes.data, payload.nbytes)
        if r is None:
            break
        reqs.append(r)
    assert len(reqs) == 64
    # the two accepted connections are distinct peers: whatever is sent on sc2 must come out of exactly one of them
    marker = np.full(300_000, 7, dtype=np.uint8)
    out = [np.zeros(300_000, dtype=np.uint8) for _ in range(2)]
    net.wait(net.isend(sc2, marker.ctypes.data, marker.nbytes))
    got = []
    for comm, buf in zip((rc1, rc2), out):
        bufs = [np.zeros(300_000, dtype=np.uint8) for _ in range(65)]
        rr = [net.irecv(comm, b.ctypes.data, b.nbytes) for b in bufs[:64]]
        # one of the two connections carries the 64 small messages, the other the marker
        first = None
        import time

        deadline = time.time() + 20
        while first is None and time.time() < deadline:
            done, size, rc = net.test(rr[0])
            if done:
                first = size
        got.append(first)
        if first == 300_000:
            np.testing.assert_array_equal(bufs[0], marker)
        else:
            assert first == 256
            for r in rr[1:]:
                assert net.wait(r) == 256
        # the requests of the other connection stay pending until close → closing must not hang
    assert sorted(got) == [256, 300_000]
    for r in reqs:
        assert net.wait(r) == 256
    for c in (sc1, sc2):
        net.close_send(c)
    for c in (rc1, rc2):
        net.close_recv(c)
    net.close_listen(lc)


def test_enable_sets_nccl_environment():
    from bagua_b200 import net as bnet

    env = {"LD_LIBRARY_PATH": "/x"}
    bnet.enable(env, nstreams=8, min_chunksize=1 << 19)
    assert env["NCCL_NET_PLUGIN"] == "bagua" and env["BAGUA_NET_NSTREAMS"] == "8" and env["BAGUA_NET_MIN_CHUNKSIZE"] == str(1 << 19)
    first = env["LD_LIBRARY_PATH"].split(":")[0]
    assert os.path.exists(os.path.join(first, "libnccl-net-bagua.so")) and env["LD_LIBRARY_PATH"].endswith("/x")


def test_span_trace_file(tmp_path):
    """BAGUA_NET_TRACE_FILE: every completed isend/irecv becomes one trace-event span with its byte count (the reference ships
    spans to Jaeger, nthread_per_socket_backend.rs:113-137).  The tracer reads its environment once per process → subprocess."""
    import json
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from bagua_b200 import net as bnet
bnet.build_if_needed() if hasattr(bnet, "build_if_needed") else None
h = bnet.PluginHandle()
handle, lc = h.listen(0)
sc = h.connect(handle)
rc = h.accept(lc)
for n in (64, 100000, 3 << 20):
    src = np.arange(n, dtype=np.uint8)
    dst = np.zeros(n, dtype=np.uint8)
    r = h.irecv(rc, dst.ctypes.data, n)
    s = h.isend(sc, src.ctypes.data, n)
    h.wait(s); h.wait(r)
    assert (src == dst).all()
h.trace_flush()
h.close_send(sc); h.close_recv(rc); h.close_listen(lc)
print("TRACE_OK")
''' % repo
    env = dict(os.environ, BAGUA_NET_TRACE_FILE=str(tmp_path / "spans"), RANK="3", NCCL_SOCKET_IFNAME="lo", BAGUA_NET_NSTREAMS="3", BAGUA_NET_MIN_CHUNKSIZE="65536")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "TRACE_OK" in r.stdout, r.stdout + r.stderr
    events = json.loads((tmp_path / "spans.3.json").read_text())
    assert sorted((e["name"], e["args"]["bytes"]) for e in events) == sorted([(k, n) for n in (64, 100000, 3 << 20) for k in ("isend", "irecv")])
    assert all(e["ph"] == "X" and e["pid"] == 3 and e["dur"] >= 0 and e["args"]["ok"] for e in events)


@pytest.mark.parametrize("family", ["AF_INET", "2", "4", "AF_INET6", "nonsense"])
def test_socket_family_filter(family):
    """NCCL_SOCKET_FAMILY (reference rust/bagua-net/src/utils.rs:33-36,101): NCCL's spelling, the numeric sa_family and 4 / 6 select
    the address family of the candidate interfaces; an unknown value means no restriction.  IPv4 loopback always exists here; IPv6
    may not, in which case the filtered list is empty and the plugin says so instead of picking an IPv4 address."""
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from bagua_b200 import net as bnet
try:
    h = bnet.PluginHandle()
    devs = h.devices()
except Exception as e:
    print("NO_DEVICE", type(e).__name__)
    raise SystemExit(0)
if not devs:
    print("NO_DEVICE empty")
    raise SystemExit(0)
handle, lc = h.listen(0)
sc = h.connect(handle)
rc = h.accept(lc)
src = np.arange(100000, dtype=np.uint8); dst = np.zeros(100000, dtype=np.uint8)
r = h.irecv(rc, dst.ctypes.data, src.size); s = h.isend(sc, src.ctypes.data, src.size)
h.wait(s); h.wait(r)
assert (src == dst).all()
h.close_send(sc); h.close_recv(rc); h.close_listen(lc)
print("ROUNDTRIP_OK", devs[0]["name"])
''' % repo
    env = dict(os.environ, NCCL_SOCKET_IFNAME="lo", NCCL_SOCKET_FAMILY=family)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    if family in ("AF_INET", "2", "4", "nonsense"):
        assert "ROUNDTRIP_OK lo" in r.stdout, r.stdout + r.stderr
    else:
        assert "ROUNDTRIP_OK lo" in r.stdout or "NO_DEVICE" in r.stdout, r.stdout + r.stderr
