r"""Elastic launcher (reference: bagua/distributed/run.py:1-639, itself a fork of ``torch.distributed.run``).

All of torchrun's flags (``--nnodes MIN:MAX``, ``--rdzv_backend/--rdzv_endpoint/--rdzv_id``, ``--max_restarts``,
``--monitor_interval``, ``--standalone`` …) plus the bagua flags of :mod:`bagua_b200.distributed.launch`.  Fault-tolerance
semantics are torch elastic's: when a worker fails or membership changes, *all* workers are restarted (up to
``--max_restarts``) with re-assigned ranks; the training script resumes from its own checkpoint (see
``examples/elastic_training``).  Symmetric-memory heaps and signal pads are created per process, so a restart starts clean.

    python -m bagua_b200.distributed.run --standalone --nnodes=1 --nproc_per_node=8 train.py
    python -m bagua_b200.distributed.run --nnodes=1:4 --nproc_per_node=8 --rdzv_id=job --rdzv_backend=c10d --rdzv_endpoint=host:29400 train.py
"""
from __future__ import annotations

import os
from argparse import ArgumentParser

from .launch import add_bagua_arguments, set_bagua_env

# helpers of torch's launcher that the reference's run.py re-implements (bagua/distributed/run.py:436-585); same objects here
from torch.distributed.run import config_from_args, determine_local_world_size, get_rdzv_endpoint, parse_min_max_nnodes, run_script_path  # noqa: E402,F401


def get_args_parser() -> ArgumentParser:
    from torch.distributed.run import get_args_parser as torch_parser

    parser = torch_parser()
    add_bagua_arguments(parser)
    if not any("--use_env" in a.option_strings or "--use-env" in a.option_strings for a in parser._actions):
        # torch ≥ 1.10 always passes LOCAL_RANK through the environment; the reference's launcher still takes the switch (run.py:424-430)
        parser.add_argument("--use_env", "--use-env", default=True, action="store_true", help="accepted for compatibility: LOCAL_RANK is always set in the environment")
    return parser


def parse_args(args=None):
    return get_args_parser().parse_args(args)


def run(args):
    from torch.distributed.run import run as torch_run

    if args.standalone:
        args.rdzv_backend = "c10d"
        args.rdzv_endpoint = "127.0.0.1:29400" if not getattr(args, "rdzv_endpoint", "") else args.rdzv_endpoint
        args.rdzv_id = args.rdzv_id if getattr(args, "rdzv_id", "none") not in ("", "none") else "bagua_standalone"
    master_addr = getattr(args, "master_addr", "127.0.0.1") or "127.0.0.1"
    if not hasattr(args, "master_addr") or args.master_addr is None:
        args.master_addr = master_addr
    # the workers inherit the launcher's environment
    env = {}
    set_bagua_env(args, env)
    os.environ.update(env)
    torch_run(args)


def main(args=None):
    args = parse_args(args)
    run(args)


if __name__ == "__main__":
    main()
