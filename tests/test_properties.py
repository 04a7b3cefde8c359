"""Property-based tests (hypothesis) of the host-side invariants the kernels and the scheduler rely on."""
import math

import pytest
import torch

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

DTYPES = ["f32", "f16", "bf16"]
UNIT = {"f32": 4, "f16": 2, "bf16": 2}


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.tuples(st.integers(1, 5_000_000), st.sampled_from(DTYPES)), min_size=1, max_size=40), st.integers(10, 26))
def test_bucket_split_is_a_dtype_homogeneous_ordered_partition(tensors, size_2p):
    """every tensor lands in exactly one bucket, a bucket holds one dtype, registration order is kept inside
    a dtype, and a bucket closes as soon as it reaches the size (so only the last bucket of a dtype may be smaller)."""
    from bagua_b200.define import TensorDeclaration
    from bagua_b200.service.autotune_task_manager import split_bucket_by_bucket_size

    decls = [TensorDeclaration(name=f"t{i}", num_elements=n, dtype=d) for i, (n, d) in enumerate(tensors)]
    bucket_size = 1 << size_2p
    buckets = split_bucket_by_bucket_size(decls, bucket_size)
    flat = [td["name"] for b in buckets for td in b]
    assert sorted(flat) == sorted(d["name"] for d in decls) and len(set(flat)) == len(flat)
    per_dtype = {}
    for b in buckets:
        kinds = {str(getattr(td["dtype"], "value", td["dtype"])) for td in b}
        assert len(kinds) == 1
        per_dtype.setdefault(kinds.pop(), []).append(b)
    for kind, bs in per_dtype.items():
        order = [int(td["name"][1:]) for b in bs for td in b]
        assert order == sorted(order)
        for b in bs[:-1]:
            nbytes = sum(td["num_elements"] * UNIT[kind] for td in b)
            assert nbytes >= bucket_size and nbytes - b[-1]["num_elements"] * UNIT[kind] < bucket_size


@settings(max_examples=100, deadline=None, derandomize=True)
@given(st.integers(1, 8), st.integers(1, 4000), st.floats(1e-3, 1e3), st.floats(-1e3, 1e3), st.sampled_from([torch.float32, torch.float16]))
def test_minmax_uint8_roundtrip_error_bound(n_chunks, chunk, spread, offset, dtype):
    """MinMaxUInt8 (reference bagua_kernels.cu:456-501): every chunk decodes to within one quantisation step of its input, the
    compressed size follows the 32-byte-aligned wire format, and constant chunks survive exactly."""
    from bagua_b200.ops import quant

    torch.manual_seed(chunk * 31 + n_chunks)
    if dtype == torch.float16:
        spread, offset = min(spread, 100.0), max(min(offset, 100.0), -100.0)
    x = (torch.rand(n_chunks * chunk) * spread + offset).to(dtype)
    buf = quant.torch_compress(x, n_chunks)
    assert buf.dtype == torch.uint8 and buf.numel() == quant.compressed_size(x.numel(), n_chunks)
    y = quant.torch_decompress(buf, x.numel(), n_chunks, dtype)
    for c in range(n_chunks):
        xs, ys = x[c * chunk:(c + 1) * chunk].float(), y[c * chunk:(c + 1) * chunk].float()
        step = (xs.max() - xs.min()).item() / 255.0
        tol = step * 1.01 + (abs(xs).max().item() + 1.0) * (2e-3 if dtype == torch.float16 else 2e-6) + 1e-6
        assert (xs - ys).abs().max().item() <= tol


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.integers(2, 64), st.integers(0, 500))
def test_shift_one_pairs_every_rank_with_exactly_one_partner(nranks, step):
    """shift_one peer selection (reference decentralized_full_precision_synchronous.rs:83-96) is an involution without fixed
    points for even world sizes: rank r's partner names r back, so the pairwise exchange cannot deadlock."""
    from bagua_b200.core import native

    if nranks % 2:
        nranks += 1
    f = native().PeerAverageOp.shift_one_peer
    peers = [f(r, nranks, step) for r in range(nranks)]
    assert sorted(peers) == list(range(nranks))
    assert all(peers[p] == r and p != r for r, p in enumerate(peers))


def _flatten_worker(rank, world):
    """Runs inside a one-rank process group (buckets register with the default group's scheduler)."""
    import random

    import bagua_b200 as bagua
    from bagua_b200.bucket import BaguaBucket

    bagua.init_process_group()
    rng = random.Random(7)
    for case in range(60):
        alignment = rng.choice([1, 4, 8, 64])
        shapes = [[rng.randint(1, 7) for _ in range(rng.randint(0, 3))] for _ in range(rng.randint(1, 8))]
        ts = [(torch.randn(*shp) if shp else torch.randn(())).ensure_bagua_tensor(f"c{case}p{i}", f"prop{case}") for i, shp in enumerate(shapes)]
        before = [t.clone() for t in ts]
        b = BaguaBucket(ts, f"b{case}", flatten=True, alignment=alignment)
        assert b.check_flatten()
        total = sum(t.numel() for t in ts)
        flat = b.backend_tensor
        assert flat.numel() >= total and flat.numel() % alignment == 0 and flat.numel() - total < alignment
        off = 0
        for t, ref in zip(ts, before):
            assert torch.equal(t, ref)
            assert t.data_ptr() == flat.data_ptr() + off * t.element_size()
            off += t.numel()
        flat.fill_(3.0)  # writes through the flat tensor are visible in every member
        assert all(bool((t == 3.0).all()) for t in ts)
    return True


def test_flattened_bucket_aliases_every_tensor():
    """BaguaBucket(flatten=True): tensors become views of one flat tensor (in order, without gaps), values are preserved, the
    padded length is the next multiple of the alignment and ``check_flatten`` holds — 60 random shape lists."""
    from tests.mp_utils import run_distributed

    assert all(run_distributed(_flatten_worker, world=1))


@settings(max_examples=6, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 31), st.integers(5, 16))
def test_bayesian_optimizer_stays_inside_the_declared_space(seed, rounds):
    from bagua_b200.service.bayesian_optimizer import BayesianOptimizer, BoolParam, FloatParam, IntParam

    space = {"i": IntParam(val=3, space_dimension=(1, 9)), "f": FloatParam(val=0.5, space_dimension=(0.25, 4.0)), "b": BoolParam(False)}
    opt = BayesianOptimizer(space, n_initial_points=4, seed=seed % 1000)
    for _ in range(rounds):
        p = opt.ask()
        assert isinstance(p["i"], int) and 1 <= p["i"] <= 9
        assert 0.25 <= p["f"] <= 4.0 and isinstance(p["b"], bool)
        opt.tell(p, -((p["i"] - 6) ** 2) - (math.log2(p["f"])) ** 2 + (0.5 if p["b"] else 0.0))


@settings(max_examples=120, deadline=None, derandomize=True)
@given(st.lists(st.tuples(st.lists(st.integers(1, 5), min_size=1, max_size=4), st.booleans()), min_size=1, max_size=6), st.integers(1, 8), st.integers(1, 8),
       st.sampled_from([4, 8]))
def test_optimizer_shards_consolidate_and_reshard_for_any_world_size(tensors, world_save, world_load, per):
    """Checkpoint format of the in-bucket optimizers: shards written by ``world_save`` ranks consolidate into per-parameter
    tensors in logical order (also for channels_last parameters) and re-shard exactly for ``world_load`` ranks."""
    from bagua_b200.parallel.algorithms.gradient_allreduce import consolidate_shards, shard_of
    from bagua_b200.tensor import dense_strides

    torch.manual_seed(len(tensors) * 17 + world_save)
    params, layout, off = {}, [], 0
    for i, (shape, channels_last) in enumerate(tensors):
        t = torch.randn(*shape)
        if channels_last and t.dim() == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        params[f"p{i}"] = t
        layout.append((f"p{i}", off, t.numel(), tuple(t.shape), tuple(dense_strides(t))))
        off += t.numel()
    numel = (off + per - 1) // per * per
    flat = torch.zeros(numel)
    for name, o, n, shape, strides in layout:
        flat[o:o + n] = torch.as_strided(params[name], (n,), (1,))       # memory order, as the bucket holds it

    def shards(world):
        length = ((numel // per + world - 1) // world) * per
        out = []
        for r in range(world):
            lo, hi = r * length, min((r + 1) * length, numel)
            s = torch.zeros(length)
            if hi > lo:
                s[: hi - lo] = flat[lo:hi]
            out.append((s, lo, max(lo, hi), length))
        return out

    cons = consolidate_shards([s for s, _, _, _ in shards(world_save)], numel, layout)
    for name, t in params.items():
        assert torch.equal(cons[name], t.contiguous()) and cons[name].is_contiguous()
    for s, lo, hi, length in shards(world_load):
        assert torch.equal(shard_of(cons, layout, numel, lo, hi, length), s)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.integers(1, 1 << 30), st.integers(1, 16), st.integers(12, 24), st.integers(0, 1000))
def test_net_plugin_chunk_plan_partitions_the_message(size, nstreams, min_chunk_2p, cursor):
    """csrc/net plan_chunks (both ends of a connection derive it from the message size alone): contiguous cover of [0, size), at
    most one chunk per stream, cuts on 16-byte boundaries, consecutive streams starting at the rotating cursor, and no more
    chunks than ``size / min_chunk`` allows."""
    __import__("os").environ.setdefault("NCCL_SOCKET_IFNAME", "lo")
    h = _plugin_handle()
    min_chunk = 1 << min_chunk_2p
    plan = h.plan_chunks(size, nstreams, min_chunk, cursor)
    assert 1 <= len(plan) <= nstreams and len(plan) <= max(1, size // min_chunk)
    pos = 0
    for i, (off, nbytes, stream) in enumerate(plan):
        assert off == pos and nbytes > 0 and stream == (cursor + i) % nstreams
        assert off % 16 == 0
        pos += nbytes
    assert pos == size


_handle_cache = []


def _plugin_handle():
    if not _handle_cache:
        from bagua_b200 import net

        _handle_cache.append(net.PluginHandle())
    return _handle_cache[0]


def test_bytegrad_bucket_merging_keeps_order_dtype_boundaries_and_the_size_floor():
    """``merge_small_buckets`` (bytegrad.py): consecutive suggested buckets are concatenated until each holds ``min_bytes``; tensor order
    is preserved (a merged bucket becomes ready when its last member would have), dtypes never mix, a small remainder joins its predecessor."""
    import random

    import torch

    from bagua_b200.parallel.algorithms.bytegrad import merge_small_buckets

    class T:   # what the function touches of a bagua tensor
        def __init__(self, numel, dtype, tag):
            self._t, self.tag = torch.empty(numel, dtype=dtype), tag

        def bagua_getter_closure(self):
            return self._t

    def nbytes(bucket):
        return sum(t._t.numel() * t._t.element_size() for t in bucket)

    rng = random.Random(5)
    for trial in range(200):
        dtypes = [torch.float32] * rng.randint(1, 6) + [torch.bfloat16] * rng.randint(0, 6) + [torch.float32] * rng.randint(0, 3)
        buckets, tag = [], 0
        for dt in dtypes:
            b = []
            for _ in range(rng.randint(1, 4)):
                b.append(T(rng.randint(1, 400), dt, tag))
                tag += 1
            buckets.append(b)
        floor = rng.choice([0, 64, 1000, 4000, 10 ** 6])
        merged = merge_small_buckets(buckets, floor)
        assert [t.tag for b in merged for t in b] == list(range(tag)), "order / completeness"
        assert all(len({t._t.dtype for t in b}) == 1 for b in merged), "mixed dtypes in one bucket"
        if floor == 0:
            assert merged == buckets
            continue
        # within one run of equal dtype only the LAST bucket may be under the floor
        runs, cur = [], [merged[0]]
        for b in merged[1:]:
            if b[0]._t.dtype == cur[-1][0]._t.dtype:   # runs are adjacent buckets of one dtype
                cur.append(b)
            else:
                runs.append(cur)
                cur = [b]
        runs.append(cur)
        for run in runs:
            assert all(nbytes(b) >= floor for b in run[:-1]), (trial, [nbytes(b) for b in run], floor)
        assert len(merged) <= len(buckets)
