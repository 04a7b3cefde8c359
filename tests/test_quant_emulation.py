"""The scalar MinMaxUInt8 math of the CUDA kernels (csrc/quant.cuh), compiled for the HOST (tests/cpp/quant_emulation.cpp: every
intrinsic the header uses is one IEEE round-to-nearest operation, so plain C++ without FMA contraction reproduces the device bit for
bit) and checked against the python oracle (bagua_b200.ops.quant.torch_*_chunk, itself the formula of the reference's
tests/internal/compressor.py:4-33).  This is how the CPU suite sees the numerics of the decode that multiplies by the chunk's
reciprocal instead of dividing — a change made without GPU access."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp, extra=()):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(tmp, "libquant_emu" + ("_div" if extra else "") + ".so")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", f"-I{REPO}/bagua_b200/csrc", *extra,
           os.path.join(REPO, "tests", "cpp", "quant_emulation.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(out)
    lib.emu_minmax_uint8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.emu_decode_16bit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    lib.emu_f32_to_ordered.argtypes, lib.emu_f32_to_ordered.restype = [ctypes.c_float], ctypes.c_uint32
    lib.emu_ordered_to_f32.argtypes, lib.emu_ordered_to_f32.restype = [ctypes.c_uint32], ctypes.c_float
    return lib


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("quant_emu"))
    return _build(tmp), _build(tmp, ("-DBAGUA_DEQUANT_IEEE_DIV",))


def _cases():
    g = torch.Generator().manual_seed(7)
    yield "normal", torch.randn(4096, generator=g) * 3
    yield "tiny range", 1.0 + torch.rand(1024, generator=g) * 1e-6
    yield "constant", torch.full((256,), 0.37)
    yield "large", torch.randn(2048, generator=g) * 1e6
    yield "small", torch.randn(2048, generator=g) * 1e-8
    yield "one sided", torch.rand(2048, generator=g) * 7 + 100
    yield "negative", -torch.rand(2048, generator=g) * 50
    yield "gradients", torch.randn(8192, generator=g) * 1e-3 * torch.randn(8192, generator=g).abs()
    yield "halves", torch.arange(-128, 128, dtype=torch.float32) * 0.5          # products ending in .5: round-half-even territory


def _run(lib, x):
    xn = x.numpy().astype(np.float32)
    q = np.zeros(xn.size, dtype=np.uint8)
    deq = np.zeros(xn.size, dtype=np.float32)
    params = np.zeros(4, dtype=np.float32)
    lib.emu_minmax_uint8(xn.ctypes.data, xn.size, float(xn.min()), float(xn.max()), q.ctypes.data, deq.ctypes.data, params.ctypes.data)
    return q, deq, params


def _ulp_diff(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia, ib = np.where(ia < 0, -(ia & 0x7fffffff), ia), np.where(ib < 0, -(ib & 0x7fffffff), ib)      # sign-magnitude → two's complement order
    return np.abs(ia - ib)


@pytest.mark.parametrize("name,x", list(_cases()), ids=[n for n, _ in _cases()])
def test_quantised_bytes_equal_the_oracle_and_the_reciprocal_decode_is_within_one_ulp(libs, name, x):
    from bagua_b200.ops import quant

    lib, lib_div = libs
    mm, q_ref = quant.torch_compress_chunk(x)
    deq_ref = quant.torch_decompress_chunk(mm, q_ref, torch.float32).numpy()
    q, deq, params = _run(lib, x)
    q_div, deq_div, params_div = _run(lib_div, x)
    assert np.array_equal(q, q_ref.numpy()) and np.array_equal(q_div, q), "wire bytes must not depend on the decode flavour and must equal the oracle"
    assert np.array_equal(params[:3], params_div[:3])
    assert np.array_equal(deq_div.view(np.uint32), deq_ref.view(np.uint32)), "the IEEE-division build is the reference's formula, bit for bit"
    assert _ulp_diff(deq, deq_ref).max() <= 1, (name, _ulp_diff(deq, deq_ref).max())
    # decode error of the round trip: half a level of the chunk's range (plus the 1e-7 the scale formula adds to the range)
    step = (float(x.max()) - float(x.min()) + 1e-7) / 255.0
    assert np.abs(deq - x.numpy()).max() <= 0.5 * step * (1 + 1e-5) + 1e-6 * np.abs(x.numpy()).max(), name


@pytest.mark.parametrize("kind,dtype", [(1, torch.bfloat16), (2, torch.float16)])
def test_decode_into_16_bit_types_matches_the_oracle_except_at_rounding_ties(libs, kind, dtype):
    """The kernels store the decoded value through from_f32<T>: a one-ulp fp32 difference can only change the 16-bit result when the
    quotient sits (within an ulp) on a rounding boundary — rare, and then the results are neighbours."""
    from bagua_b200.ops import quant

    lib, lib_div = libs
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(1 << 16, generator=g) * 2).to(dtype)
    mm, q = quant.torch_compress_chunk(x)
    want = quant.torch_decompress_chunk(mm, q, dtype)
    qn = q.numpy()
    for lb, exact in ((lib_div, True), (lib, False)):
        out = np.zeros(qn.size, dtype=np.uint16)
        lb.emu_decode_16bit(qn.ctypes.data, qn.size, float(mm[0].float()), float(mm[1].float()), kind, out.ctypes.data)
        got = torch.from_numpy(out.view(np.int16).copy()).view(dtype)
        differ = (got != want)
        if exact:
            assert not differ.any()
        else:
            assert differ.float().mean().item() < 1e-3
            assert (got.float() - want.float()).abs().max().item() <= (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10) * want.float().abs().max().item()


def test_order_preserving_float_encoding(libs):
    lib, _ = libs
    vals = [-float("inf"), -3.0e38, -1.0, -1e-30, -0.0, 0.0, 1e-30, 1.0, 3.0e38, float("inf")]
    enc = [lib.emu_f32_to_ordered(v) for v in vals]
    assert enc == sorted(enc) and len(set(enc[:4] + enc[6:])) == 8          # monotone; -0.0 and 0.0 may only differ in the sign bit
    for v, e in zip(vals, enc):
        back = lib.emu_ordered_to_f32(e)
        assert back == float(np.float32(v)) and np.signbit(back) == np.signbit(v)
    assert lib.emu_f32_to_ordered(2.5) < 0xFFFFFFFF and lib.emu_f32_to_ordered(-2.5) > 0   # identities of atomicMin / atomicMax are never hit by data
