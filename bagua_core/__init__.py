"""Drop-in for the reference's native python module ``bagua_core`` (rust/bagua-core/bagua-core-py/src/lib.rs:540-568;
shim package bagua_core/__init__.py): the same class names on top of the C++ core of this repository
(``bagua_b200._C``).  Tensors are described by raw pointers here, so ``BaguaTensorPy`` takes a torch tensor and extracts
them once; communicators are the (torch ProcessGroup, stream) pairs of ``bagua_b200.communication``."""
from __future__ import annotations

from typing import List, Optional

import torch

from bagua_b200.core import dtype_code, native

__all__ = ["BaguaCommBackendPy", "BaguaTensorPy", "BaguaBucketPy", "BaguaSingleCommunicatorPy", "show_version", "install_deps"]


def show_version() -> str:
    v = native().show_version()
    print(v)
    return v


def install_deps():
    """The reference downloads an NCCL tarball here (bagua_core/bagua_install_deps.py); nothing to install: the native core
    is built in-tree from source (``python -m bagua_b200._build``)."""
    return None


class BaguaTensorPy:
    def __init__(self, name: str, torch_tensor: torch.Tensor):
        self.torch_tensor = torch_tensor
        self.inner = native().Tensor(name, torch_tensor.data_ptr(), torch_tensor.numel(), dtype_code(torch_tensor.dtype),
                                     torch_tensor.device.index if torch_tensor.is_cuda else -1)

    def name(self) -> str:
        return self.inner.name()

    def data_ptr(self) -> int:
        return self.inner.data_ptr()

    def device_id(self) -> int:
        return self.inner.device_id()

    def num_elements(self) -> int:
        return self.inner.num_elements()

    def num_elements_allocated(self) -> int:
        return self.inner.num_elements()

    def dtype(self) -> str:
        return {0: "f32", 1: "f16", 2: "u8", 3: "i64", 4: "bf16"}[self.inner.dtype()]

    def compress(self, method: str, n_chunks: int, target_chunk: int = -1) -> "BaguaTensorPy":
        from bagua_b200.ops import quant

        assert method == "MinMaxUInt8"
        return BaguaTensorPy(self.name() + "_compressed", quant.compress(self.torch_tensor.contiguous().view(-1), n_chunks, target_chunk))

    def decompress_from(self, method: str, n_chunks: int, compressed: "BaguaTensorPy"):
        from bagua_b200.ops import quant

        assert method == "MinMaxUInt8"
        quant.decompress(compressed.torch_tensor, self.torch_tensor.view(-1), n_chunks)

    def to_numpy_f32(self):
        return self.torch_tensor.detach().float().cpu().numpy()

    def to_numpy_u8(self):
        return self.torch_tensor.detach().to(torch.uint8).cpu().numpy()


class BaguaBucketPy:
    def __init__(self, name: str, tensors: List[BaguaTensorPy]):
        self._tensors = tensors
        self.inner = native().Bucket(name, [t.inner for t in tensors])

    def tensors(self) -> List[BaguaTensorPy]:
        return self._tensors

    def append_python_op(self, op):
        self.inner.append_python_op(op, "python")

    # -- the reference's low-level op builders (bagua-core-py/src/lib.rs:431-520), composed from python ops here: users of
    #    ``bagua_core`` work with loose tensors, not with the symmetric bucket arena the fused kernels need ----------------
    def _flat_or_gather(self):
        ts = [t.torch_tensor for t in self._tensors]
        flat = torch.cat([t.reshape(-1) for t in ts])

        def scatter_back():
            off = 0
            with torch.no_grad():
                for t in ts:
                    t.copy_(flat[off: off + t.numel()].view_as(t))
                    off += t.numel()

        return flat, scatter_back

    def append_centralized_synchronous_op(self, communicator_internode=None, communicator_intranode=None, hierarchical: bool = False,
                                          average: bool = True, scattergather: bool = False, compression: Optional[str] = None):
        """All-reduce (optionally MinMaxUInt8-compressed) of the bucket's tensors over the default group."""
        from bagua_b200 import communication as comm_mod
        from bagua_b200.bucket import _torch_allreduce
        from bagua_b200.ops import quant

        pg = comm_mod._get_default_group()

        def run(_name: str):
            if compression is not None:
                assert compression == "MinMaxUInt8", f"unknown compression {compression}"
                quant.bytegrad_allreduce_fallback(self, pg, average)
                return
            flat, scatter_back = self._flat_or_gather()
            _torch_allreduce(flat, pg, average, hierarchical)
            scatter_back()

        self.inner.append_python_op(run, "centralized_synchronous")

    def append_decentralized_synchronous_op(self, communicator_internode=None, communicator_intranode=None, hierarchical: bool = True,
                                            peer_selection_mode: str = "all", peer_weight: Optional["BaguaTensorPy"] = None):
        """``peer_weight`` ← average of the bucket over all ranks (``all``) or with this step's partner (``shift_one``)."""
        import torch.distributed as dist

        from bagua_b200 import communication as comm_mod
        from bagua_b200.core import native as _native

        pg = comm_mod._get_default_group()
        assert peer_weight is not None, "decentralized op needs a peer_weight tensor"
        state = {"step": 0}

        def run(_name: str):
            flat, _ = self._flat_or_gather()
            out = peer_weight.torch_tensor.view(-1)
            n, r = pg.size(), pg.rank()
            if peer_selection_mode == "all":
                out.copy_(flat)
                dist.all_reduce(out, group=pg.torch_group)
                out.div_(n)
            elif peer_selection_mode == "shift_one":
                peer = _native().PeerAverageOp.shift_one_peer(r, n, state["step"])
                recv = torch.empty_like(flat)
                reqs = [dist.isend(flat, pg.ranks[peer], group=pg.torch_group), dist.irecv(recv, pg.ranks[peer], group=pg.torch_group)]
                for q in reqs:
                    q.wait()
                out.copy_((flat + recv) / 2)
                state["step"] += 1
            else:
                raise ValueError(f"unsupported peer_selection_mode {peer_selection_mode}")

        self.inner.append_python_op(run, "decentralized_synchronous")

    def append_low_precision_decentralized_synchronous_op(self, communicator_internode=None, communicator_intranode=None, hierarchical: bool = True,
                                                           peer_selection_mode: str = "ring", compression: str = "MinMaxUInt8", weight: Optional["BaguaTensorPy"] = None,
                                                           left_peer_weight: Optional["BaguaTensorPy"] = None, right_peer_weight: Optional["BaguaTensorPy"] = None):
        """Ring exchange of MinMaxUInt8-compressed weight differences with the left and right neighbour (bagua-core-py/src/lib.rs:479-502;
        comm_ops/decentralized_low_precision_synchronous.rs): the bucket holds the freshly stepped weights, ``weight`` / ``left_peer_weight`` /
        ``right_peer_weight`` are the replicas the algorithm keeps between steps."""
        from bagua_b200 import communication as comm_mod
        from bagua_b200.ops import quant

        assert compression == "MinMaxUInt8", f"unknown compression {compression}"
        assert weight is not None and left_peer_weight is not None and right_peer_weight is not None
        pg = comm_mod._get_default_group()
        self.inner.append_python_op(lambda _name: quant.low_precision_ring_fallback(self, pg, weight.torch_tensor, left_peer_weight.torch_tensor,
                                                                                    right_peer_weight.torch_tensor), "low_precision_decentralized_synchronous")

    def append_decentralized_asynchronous_op(self, communicator_internode=None, communicator_intranode=None, peer_selection_mode: str = "all", torch_stream: int = 0):
        """Model averaging that runs beside training (bagua-core-py/src/lib.rs:504-519): every execution snapshots the bucket, averages the
        snapshot over all ranks and adds ``average − snapshot`` to the live tensors under the weight lock.  Returns the op object with the
        reference's ``lock_weight / unlock_weight / abort / reset / get_status``."""
        assert peer_selection_mode == "all", "only peer_selection_mode='all' is supported (as in the reference)"
        op = DecentralizedFullPrecisionAsynchronousPy(self)
        self.inner.append_python_op(op.run, "decentralized_asynchronous")
        return op

    def print_ops(self):
        print(self.inner.print_ops())

    def clear_ops(self):
        self.inner.clear_ops()

    def ready_for_comm(self) -> bool:
        return self.inner.ready_for_comm()

    def reset_comm_ready(self):
        self.inner.reset_comm_ready()


class DecentralizedFullPrecisionAsynchronousPy:
    """Handle of an asynchronous model-average op on loose tensors (reference: bagua-core-py/src/lib.rs:394-428 over
    comm_ops/decentralized_full_precision_asynchronous.rs)."""

    def __init__(self, bucket: "BaguaBucketPy"):
        import threading

        self._bucket = bucket
        self._lock = threading.Lock()
        self._aborted = False

    def lock_weight(self):
        self._lock.acquire()

    def unlock_weight(self):
        if self._lock.locked():
            self._lock.release()

    def abort(self):
        self._aborted = True

    def reset(self):
        self._aborted = False

    def get_status(self) -> bool:
        """True while the op takes part in averaging rounds."""
        return not self._aborted

    def run(self, _name: str = ""):
        import torch.distributed as dist

        from bagua_b200 import communication as comm_mod

        pg = comm_mod._get_default_group()
        # every rank must agree on whether this round happens: MIN over "I am still running" (the reference negotiates the same way)
        go = torch.tensor([0 if self._aborted else 1], dtype=torch.int32)
        dist.all_reduce(go, op=dist.ReduceOp.MIN, group=pg.torch_group)
        if int(go.item()) == 0:
            return
        snapshot, _ = self._bucket._flat_or_gather()
        avg = snapshot.clone()
        dist.all_reduce(avg, group=pg.torch_group)
        avg.div_(pg.size())
        with self._lock, torch.no_grad():
            off = 0
            for t in (bt.torch_tensor for bt in self._bucket._tensors):
                n = t.numel()
                t.add_((avg[off: off + n] - snapshot[off: off + n]).view_as(t))
                off += n


class BaguaCommBackendPy:
    def __init__(self, schedule_channel_cap: int, device_id: int, comm_stream: int = 0):
        self.inner = native().Backend(schedule_channel_cap, device_id, comm_stream, 300.0)

    def register_ordered_buckets(self, buckets: List[BaguaBucketPy]):
        self.inner.register_ordered_buckets([b.inner for b in buckets])

    def mark_communication_ready(self, tensor: BaguaTensorPy, ready_cuda_event_ptr: int = 0):
        self.inner.mark_communication_ready(tensor.inner, ready_cuda_event_ptr)

    def wait_pending_comm_ops(self) -> int:
        return self.inner.wait_pending_comm_ops(0, True)


class BaguaSingleCommunicatorPy:
    """(rank, nranks, device, stream) communicator with the reference's collective method names, backed by
    ``bagua_b200.communication.Communicator`` over the default torch process group."""

    def __init__(self, rank: int, nranks: int, device_id: int, stream_ptr: int = 0, nccl_unique_id_str: Optional[str] = None):
        from bagua_b200 import communication as comm_mod

        self._comm = comm_mod._get_default_group().get_global_communicator()
        assert self._comm.nranks() == nranks and self._comm.rank() == rank, "communicator shape must match the default process group"
        self._device_id = device_id

    @staticmethod
    def generate_nccl_unique_id_str() -> str:
        """A fresh ``ncclUniqueId`` (128 bytes, base64) as the reference returns (communicators/mod.rs:244-258).  The engine does not
        consume it — groups rendezvous through torch.distributed and symmetric memory — but user code that creates its own NCCL
        communicator can.  Without an NCCL-enabled torch build: 128 random bytes (still unique per call)."""
        import base64

        try:
            import torch.cuda.nccl as nccl

            raw = nccl.unique_id()
        except Exception:  # noqa: BLE001
            import os

            raw = os.urandom(128)
        return base64.b64encode(raw).decode("ascii")

    def __getattr__(self, name):
        return getattr(self._comm, name)

    def device_id(self) -> int:
        return self._device_id
