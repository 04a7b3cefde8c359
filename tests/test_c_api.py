"""The extern "C" ABI of the native core (reference: rust/bagua-core/bagua-core-c) driven through ctypes on the CPU backend."""
import ctypes
import os

import bagua_b200
from bagua_b200.core import native


def test_c_abi_roundtrip():
    native()
    lib = ctypes.CDLL(os.path.join(os.path.dirname(bagua_b200.__file__), "_C.so"))
    lib.bagua_last_error.restype = ctypes.c_char_p
    lib.bagua_version.restype = ctypes.c_char_p
    lib.bagua_tensor_c_create.restype = ctypes.c_void_p
    lib.bagua_tensor_c_create.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    lib.bagua_bucket_c_create.restype = ctypes.c_void_p
    lib.bagua_bucket_c_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_char_p]
    lib.bagua_comm_backend_c_create.restype = ctypes.c_void_p
    lib.bagua_comm_backend_c_create.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, ctypes.c_double]
    CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_void_p)
    lib.bagua_bucket_c_append_callback_op.argtypes = [ctypes.c_void_p, CB, ctypes.c_void_p]
    lib.bagua_comm_backend_c_register_ordered_buckets.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    lib.bagua_comm_backend_c_mark_communication_ready.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    lib.bagua_comm_backend_c_wait_pending_comm_ops.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
    assert b"bagua_b200" in lib.bagua_version()
    assert lib.bagua_peer_comm_c_signal_pad_bytes() > 0

    t0 = lib.bagua_tensor_c_create(b"a", 0x1000, 8, 0, -1)
    t1 = lib.bagua_tensor_c_create(b"b", 0x1020, 8, 0, -1)
    assert t0 and t1
    assert not lib.bagua_tensor_c_create(b"bad", 0x1, 1, 99, -1) and b"dtype" in lib.bagua_last_error()
    arr = (ctypes.c_void_p * 2)(t0, t1)
    bucket = lib.bagua_bucket_c_create(arr, 2, b"bucket0")
    assert bucket
    seen = []
    cb = CB(lambda name, _ud: seen.append(name.decode()))
    assert lib.bagua_bucket_c_append_callback_op(bucket, cb, None) == 0
    be = lib.bagua_comm_backend_c_create(8, -1, 0, 30.0)
    barr = (ctypes.c_void_p * 1)(bucket)
    assert lib.bagua_comm_backend_c_register_ordered_buckets(be, barr, 1) == 0
    for _ in range(2):
        assert lib.bagua_comm_backend_c_mark_communication_ready(be, t0, 0) == 0
        assert lib.bagua_comm_backend_c_mark_communication_ready(be, t1, 0) == 0
        assert lib.bagua_comm_backend_c_wait_pending_comm_ops(be, 0, 1) == 0
    assert seen == ["bucket0", "bucket0"]
    # registering the same bucket twice → duplicate tensor names → error code + message
    barr2 = (ctypes.c_void_p * 2)(bucket, bucket)
    assert lib.bagua_comm_backend_c_register_ordered_buckets(be, barr2, 2) != 0 and b"duplicated" in lib.bagua_last_error()
    for h, fn in ((be, lib.bagua_comm_backend_c_destroy), (bucket, lib.bagua_bucket_c_destroy), (t0, lib.bagua_tensor_c_destroy), (t1, lib.bagua_tensor_c_destroy)):
        p = ctypes.c_void_p(h)
        fn(ctypes.byref(p))
        assert not p.value
