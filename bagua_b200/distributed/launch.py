r"""Static multi-process launcher (reference: bagua/distributed/launch.py:1-343).

One worker process per GPU; exports the rendezvous variables (``MASTER_ADDR/PORT, WORLD_SIZE, RANK, LOCAL_RANK,
LOCAL_WORLD_SIZE, NODE_RANK``) plus the ``BAGUA_*`` knobs derived from the command line; if any worker exits non-zero all the
others are terminated; SIGINT/SIGTERM are forwarded.

Single node, 8 GPUs::

    python -m bagua_b200.distributed.launch --nproc_per_node=8 train.py --arg1 ...

Multi node (run on every node)::

    python -m bagua_b200.distributed.launch --nproc_per_node=8 --nnodes=2 --node_rank=<0|1> --master_addr=10.0.0.1 --master_port=1234 train.py ...
"""
from __future__ import annotations

import os
import signal
import subprocess
import sys
import time
from argparse import REMAINDER, ArgumentParser
from typing import List


def add_bagua_arguments(parser: ArgumentParser):
    parser.add_argument("--bagua_service_port", type=int, default=29501, help="port of the autotune service on the master node")
    parser.add_argument("--set_additional_flag", default=False, action="store_true", help="pass --local_rank to the script instead of only exporting LOCAL_RANK")
    parser.add_argument("--autotune_level", type=int, default=0, help="0 = off, 1 = tune bucket size / kernel variant online")
    parser.add_argument("--is_output_autotune_log", action="store_true", default=False)
    parser.add_argument("--report_metrics", action="store_true", default=False)
    parser.add_argument("--autotune_max_samples", type=int, default=60)
    parser.add_argument("--autotune_sampling_confidence_time", type=float, default=5.0)
    parser.add_argument("--autotune_warmup_time", type=float, default=30.0)
    parser.add_argument("--default_bucket_size", type=int, default=10 * 1024 ** 2, help="bucket size in bytes before autotune")
    parser.add_argument("--enable_bagua_net", action="store_true", default=False,
                        help="load the multi-stream TCP NCCL network plugin (libnccl-net-bagua.so) in the workers; only inter-node traffic uses it")
    parser.add_argument("--host_list", type=str, default="", help="(baguarun) comma separated host list")
    parser.add_argument("--ssh_port", type=int, default=22, help="(baguarun) ssh port")


def parse_args(argv=None):
    parser = ArgumentParser(description="bagua_b200 distributed training launcher: spawns one process per GPU")
    parser.add_argument("--nnodes", type=int, default=1)
    parser.add_argument("--node_rank", type=int, default=0)
    parser.add_argument("--nproc_per_node", type=int, default=1)
    parser.add_argument("--master_addr", default="127.0.0.1", type=str)
    parser.add_argument("--master_port", default=29500, type=int)
    parser.add_argument("-m", "--module", default=False, action="store_true", help="treat the script as a python module (python -m)")
    parser.add_argument("--no_python", default=False, action="store_true", help="run the script directly, without the python interpreter")
    parser.add_argument("--logdir", default=None, type=str, help="write each worker's stdout/stderr to <logdir>/node_<n>_local_rank_<r>_{stdout,stderr}")
    add_bagua_arguments(parser)
    parser.add_argument("training_script", type=str)
    parser.add_argument("training_script_args", nargs=REMAINDER)
    return parser.parse_args(argv)


def set_bagua_env(args, current_env: dict):
    """CLI flags → ``BAGUA_*`` environment (reference launch.py:157-179)."""
    current_env["BAGUA_SERVICE_PORT"] = str(args.bagua_service_port)
    current_env["BAGUA_DEFAULT_BUCKET_SIZE"] = str(args.default_bucket_size)
    current_env["BAGUA_AUTOTUNE"] = str(args.autotune_level)
    current_env["BAGUA_IS_OUTPUT_AUTOTUNE_LOG"] = str(int(args.is_output_autotune_log))
    current_env["BAGUA_REPORT_METRICS"] = str(int(args.report_metrics))
    current_env["BAGUA_AUTOTUNE_MAX_SAMPLES"] = str(args.autotune_max_samples)
    current_env["BAGUA_AUTOTUNE_SAMPLING_CONFIDENCE_TIME_S"] = str(args.autotune_sampling_confidence_time)
    current_env["BAGUA_AUTOTUNE_WARMUP_TIME_S"] = str(args.autotune_warmup_time)
    if args.autotune_level > 0:
        current_env["AUTO_TUNE_SERVER_ADDR"] = f"{args.master_addr}:{args.bagua_service_port}"
    if getattr(args, "enable_bagua_net", False):
        from ..net import enable as enable_net_plugin

        enable_net_plugin(current_env)  # NCCL_NET_PLUGIN=bagua + LD_LIBRARY_PATH (reference launch.py:102-107 does the same for libnccl-net.so)


def _die_with_parent():
    """``preexec_fn`` of the workers (Linux): if the launcher is killed outright (SIGKILL, OOM), the kernel sends SIGTERM to the
    worker — no orphaned ranks spinning in a collective.  The reference's launcher leaves them behind in that case."""
    try:
        import ctypes

        ctypes.CDLL("libc.so.6", use_errno=True).prctl(1, int(signal.SIGTERM))  # PR_SET_PDEATHSIG
    except Exception:  # noqa: BLE001 - best effort, e.g. non-glibc systems
        pass


def _worker_cmd(args, local_rank: int) -> List[str]:
    cmd: List[str] = []
    if not args.no_python:
        cmd = [sys.executable, "-u"]
        if args.module:
            cmd.append("-m")
    elif args.module:
        raise ValueError("Don't use both the '--no_python' flag and the '--module' flag at the same time.")
    cmd.append(args.training_script)
    if args.set_additional_flag:
        cmd.append(f"--local_rank={local_rank}")
    cmd.extend(args.training_script_args)
    return cmd


def main(argv=None) -> int:
    args = parse_args(argv)
    world_size = args.nproc_per_node * args.nnodes
    base_env = os.environ.copy()
    base_env["MASTER_ADDR"] = args.master_addr
    base_env["MASTER_PORT"] = str(args.master_port)
    base_env["WORLD_SIZE"] = str(world_size)
    base_env["NODE_RANK"] = str(args.node_rank)
    base_env["LOCAL_WORLD_SIZE"] = str(args.nproc_per_node)
    set_bagua_env(args, base_env)
    if "OMP_NUM_THREADS" not in os.environ and args.nproc_per_node > 1:
        base_env["OMP_NUM_THREADS"] = "1"
    if args.logdir:
        os.makedirs(args.logdir, exist_ok=True)

    procs: List[subprocess.Popen] = []
    files = []
    for local_rank in range(args.nproc_per_node):
        env = dict(base_env)
        env["RANK"] = str(args.nproc_per_node * args.node_rank + local_rank)
        env["LOCAL_RANK"] = str(local_rank)
        out = err = None
        if args.logdir:
            prefix = os.path.join(args.logdir, f"node_{args.node_rank}_local_rank_{local_rank}")
            out, err = open(prefix + "_stdout", "w"), open(prefix + "_stderr", "w")
            files += [out, err]
        procs.append(subprocess.Popen(_worker_cmd(args, local_rank), env=env, stdout=out, stderr=err, preexec_fn=_die_with_parent))

    def forward(signum, _frame):
        for p in procs:
            if p.poll() is None:
                p.send_signal(signum)

    signal.signal(signal.SIGINT, forward)
    signal.signal(signal.SIGTERM, forward)

    rc = 0
    alive = set(range(len(procs)))
    try:
        while alive:
            for i in list(alive):
                ret = procs[i].poll()
                if ret is None:
                    continue
                alive.discard(i)
                if ret != 0:
                    rc = ret
                    for j in alive:  # one failure takes the job down (reference launch.py:283-300)
                        procs[j].terminate()
                    deadline = time.time() + 30
                    for j in alive:
                        try:
                            procs[j].wait(timeout=max(0.1, deadline - time.time()))
                        except subprocess.TimeoutExpired:
                            procs[j].kill()
                    alive.clear()
                    break
            time.sleep(0.05)
    finally:
        for f in files:
            f.close()
    if rc != 0:
        raise subprocess.CalledProcessError(returncode=rc, cmd=_worker_cmd(args, 0))
    return 0


if __name__ == "__main__":
    sys.exit(main())
