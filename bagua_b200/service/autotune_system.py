r"""Offline tuner of system-level (NCCL) knobs (reference: bagua/service/autotune_system.py:1-169).

Repeatedly launches ``bagua_sys_perf`` under different values of ``NCCL_MIN_NCHANNELS / NCCL_SOCKET_NTHREADS /
NCCL_NSOCKS_PERTHREAD / NCCL_BUFFSIZE`` (relevant for the multi-node NCCL legs) plus this framework's
``BAGUA_COMM_BLOCKS`` (CTAs of the NVSwitch kernels) and keeps the best; a Bayesian optimiser proposes the next point."""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys

from .bayesian_optimizer import BayesianOptimizer, IntParam


def sysperf(host_list: str, nproc_per_node: int, ssh_port: int, env: dict = {}, model: str = "vgg16", extra_args=(), master_port: int = 0) -> float:
    """One measurement: total img/s printed by ``bagua_sys_perf`` under ``env`` (0.0 when the run failed)."""
    full_env = dict(os.environ)
    full_env.update({k: str(v) for k, v in env.items()})
    if host_list:
        cmd = [sys.executable, "-m", "bagua_b200.script.baguarun", "--host_list", host_list, "--ssh_port", str(ssh_port), "--nproc_per_node", str(nproc_per_node)]
        for k in env:
            cmd += ["-x", k]
        if master_port:
            cmd += ["--master_port", str(master_port)]
        cmd += ["-m", "bagua_b200.script.bagua_sys_perf", "--model", model, *extra_args]
    else:
        cmd = [sys.executable, "-m", "bagua_b200.distributed.launch", f"--nproc_per_node={nproc_per_node}"]
        if master_port:
            cmd.append(f"--master_port={master_port}")
        cmd += ["-m", "bagua_b200.script.bagua_sys_perf", "--model", model, *extra_args]
    out = subprocess.run(cmd, env=full_env, capture_output=True, text=True).stdout
    m = re.search(r"Total img/sec on (\d+) (\S+)\(s\): (\d*\.\d+|\d+)", out)
    return float(m.group(3)) if m else 0.0


def autotune_system_hyperparameters(host_list: str, nproc_per_node: int, ssh_port: int, max_samples: int = 100, model: str = "vgg16", extra_args=(),
                                    port_fn=None):
    """Offline tuner of the NCCL system knobs (``NCCL_MIN_NCHANNELS``, ``NCCL_SOCKET_NTHREADS``, ``NCCL_NSOCKS_PERTHREAD``,
    ``NCCL_BUFFSIZE``): launches ``bagua_sys_perf --model <model>`` over ``baguarun`` for every sample, reads the throughput it
    prints and returns the best setting (reference service/autotune_system.py:92-169)."""
    optim = BayesianOptimizer(
        {
            "NCCL_MIN_NCHANNELS": IntParam(0, (0, 12)),
            "NCCL_SOCKET_NTHREADS": IntParam(0, (0, 8)),
            "NCCL_NSOCKS_PERTHREAD": IntParam(0, (0, 8)),
            "nccl_buffsize_2p": IntParam(0, (0, 26)),
            "BAGUA_COMM_BLOCKS": IntParam(0, (0, 64)),
        },
        n_initial_points=10,
    )
    best = (None, float("-inf"))
    param = optim.ask()
    for _ in range(max_samples):
        env = {k: v for k, v in param.items() if k != "nccl_buffsize_2p" and v > 0}
        if param["nccl_buffsize_2p"] > 0:
            env["NCCL_BUFFSIZE"] = 2 ** param["nccl_buffsize_2p"]
        score = sysperf(host_list, nproc_per_node, ssh_port, env, model, extra_args, port_fn() if port_fn else 0)
        if score > best[1]:
            best = (env, score)
        optim.tell(param, score)
        param = optim.ask()
    return best


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--host_list", default="")
    p.add_argument("--nproc_per_node", type=int, default=1)
    p.add_argument("--ssh_port", type=int, default=22)
    p.add_argument("--max_samples", type=int, default=20)
    p.add_argument("--model", default="vgg16")
    a = p.parse_args()
    print(autotune_system_hyperparameters(a.host_list, a.nproc_per_node, a.ssh_port, a.max_samples, a.model))
