"""Spawn-N-local-processes harness (same idea as the reference's tests/internal/multi_process.py:9-53): fabricate the
launcher environment, run ``fn(rank, world, *args)`` in every process, collect the returned (picklable) results."""
from __future__ import annotations

import os
import socket
import traceback
from contextlib import closing

import torch.multiprocessing as mp


_used_ports = set()


def free_port() -> int:
    """A port nobody listens on right now and that this test session has not handed out before (TIME_WAIT / reuse races
    between consecutive spawns showed up as EADDRINUSE on the GPU box)."""
    for _ in range(64):
        with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        if port not in _used_ports and port > 20000:
            _used_ports.add(port)
            return port
    raise RuntimeError("no free port")


def _portable(obj):
    """Tensors travel by value (numpy) so the parent never depends on a child's shared-memory handles."""
    import torch

    if isinstance(obj, torch.Tensor):
        t = obj.detach().cpu()
        return ("__tensor__", t.float().numpy() if t.dtype in (torch.bfloat16, torch.float16) else t.numpy(), str(t.dtype))
    if isinstance(obj, (list, tuple)):
        return type(obj)(_portable(o) for o in obj)
    if isinstance(obj, dict):
        return {k: _portable(v) for k, v in obj.items()}
    return obj


def _restore(obj):
    import torch

    if isinstance(obj, tuple) and len(obj) == 3 and obj[0] == "__tensor__":
        return torch.from_numpy(obj[1]).to(getattr(torch, obj[2].split(".")[1]))
    if isinstance(obj, (list, tuple)):
        return type(obj)(_restore(o) for o in obj)
    if isinstance(obj, dict):
        return {k: _restore(v) for k, v in obj.items()}
    return obj


def _entry(rank, world, port, fn, args, use_cuda, extra_env, queue):
    try:
        os.environ.update(
            RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
            MASTER_PORT=str(port), BAGUA_PEER_TIMEOUT_S="20", BAGUA_COMM_TIMEOUT_S="120",
        )
        os.environ.update(extra_env or {})
        if not use_cuda:
            os.environ["BAGUA_FORCE_CPU"] = "1"
            os.environ["CUDA_VISIBLE_DEVICES"] = ""
        else:
            import torch

            torch.cuda.set_device(rank)
        out = fn(rank, world, *args)
        queue.put((rank, "ok", _portable(out)))
    except Exception:  # noqa: BLE001
        queue.put((rank, "error", traceback.format_exc()))


def run_distributed(fn, world: int = 2, args=(), use_cuda: bool = False, timeout: float = 240.0, extra_env=None):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, use_cuda, extra_env, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    errors = []
    try:
        for _ in range(world):
            rank, status, payload = queue.get(timeout=timeout)
            if status == "ok":
                results[rank] = _restore(payload)
            else:
                errors.append(f"rank {rank}:\n{payload}")
    except Exception as e:  # queue.Empty → timeout
        errors.append(f"timeout waiting for workers: {e!r}")
    for p in procs:
        p.join(timeout=20)
        if p.is_alive():
            p.kill()
            errors.append(f"process {p.pid} had to be killed")
    if errors:
        raise AssertionError("\n".join(errors))
    return [results[r] for r in range(world)]


def _descendants(pid: int):
    """Every live process whose ancestor chain contains ``pid`` (torch elastic puts its workers into their own sessions, so a
    process-group kill alone does not reach them)."""
    children = {}
    for entry in os.listdir("/proc"):
        if not entry.isdigit():
            continue
        try:
            with open(f"/proc/{entry}/stat") as f:
                fields = f.read().rsplit(")", 1)[1].split()
            children.setdefault(int(fields[1]), []).append(int(entry))
        except (OSError, IndexError, ValueError):
            continue
    out, stack = [], [pid]
    while stack:
        for c in children.get(stack.pop(), []):
            out.append(c)
            stack.append(c)
    return out


def kill_tree(pid: int):
    import signal

    victims = _descendants(pid)
    for target in [pid] + victims:
        try:
            os.kill(target, signal.SIGKILL)
        except (ProcessLookupError, PermissionError):
            pass
    try:
        os.killpg(pid, signal.SIGKILL)
    except (ProcessLookupError, PermissionError):
        pass


def run_in_session(cmd, timeout: float, **popen_kw):
    """``subprocess.run`` for launcher commands: the command gets its own session and its whole process tree is killed on
    timeout (and swept after a normal exit), so a hung worker can never outlive the test.  Output goes through temporary
    files, not pipes: an orphan that keeps a pipe open must not be able to block the test either."""
    import subprocess
    import tempfile

    with tempfile.TemporaryFile("w+") as out, tempfile.TemporaryFile("w+") as err:
        proc = subprocess.Popen(cmd, stdout=out, stderr=err, text=True, start_new_session=True, **popen_kw)
        timed_out = False
        try:
            proc.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            timed_out = True
        finally:
            kill_tree(proc.pid)
            try:
                proc.wait(timeout=10)
            except subprocess.TimeoutExpired:
                pass
        out.seek(0)
        err.seek(0)
        stdout, stderr = out.read(), err.read()
    if timed_out:
        raise subprocess.TimeoutExpired(cmd, timeout, output=stdout, stderr=stderr)
    return subprocess.CompletedProcess(cmd, proc.returncode, stdout, stderr)
