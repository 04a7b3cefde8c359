r"""``baguarun``: start the launcher on several hosts over ssh (reference: bagua/script/baguarun.py:1-227, parallel-ssh based).

    baguarun --host_list host1,host2 --ssh_port 22 --nproc_per_node 8 [-x ENV_NAME ...] train.py --arg ...

Every host runs ``python -m bagua_b200.distributed.launch --nnodes N --node_rank i --master_addr host1 ...``; output is
streamed with a ``[host]`` prefix; the exit status is the first non-zero one."""
from __future__ import annotations

import argparse
import os
import shlex
import subprocess
import sys
import threading
import time
from typing import List


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="launch a bagua_b200 job on several hosts over ssh")
    p.add_argument("--host_list", type=str, default=None,
                   help="comma separated hosts, first = master; either plain (`a,b`, ssh port from --ssh_port / $BAGUA_SSH_PORT) or with the ssh port inline "
                        "(`a:22,b:8022`). Default: $BAGUA_NODE_DOMAIN_NAMES")
    p.add_argument("--ssh_port", type=int, default=None)
    p.add_argument("--nproc_per_node", type=int, default=1)
    p.add_argument("--master_port", type=int, default=29500)
    p.add_argument("--bagua_service_port", type=int, default=None, help="forwarded to the launcher on every host")
    p.add_argument("--no_python", action="store_true", default=False, help="forwarded to the launcher: the script is an executable, not a python file")
    p.add_argument("--enable_bagua_net", action="store_true", default=False, help="forwarded to the launcher: NCCL loads the Bagua-Net plugin")
    p.add_argument("-x", dest="export_env", action="append", default=[], help="environment variable to forward to every host: NAME (current value) or NAME=VALUE")
    p.add_argument("--dry_run", action="store_true", help="print the per-host commands instead of running them")
    p.add_argument("launch_args", nargs=argparse.REMAINDER, help="[launcher flags] script [script args]")
    args = p.parse_args(argv)
    if args.host_list is None:
        args.host_list = os.environ.get("BAGUA_NODE_DOMAIN_NAMES", "")
    if args.ssh_port is None:
        args.ssh_port = int(os.environ.get("BAGUA_SSH_PORT", 22))
    return args


def host_pairs(args) -> List[tuple]:
    """``[(host, ssh_port)]`` from either form of ``--host_list`` (reference baguarun.py:176-190)."""
    pairs = []
    for item in (h.strip() for h in args.host_list.split(",")):
        if not item:
            continue
        host, sep, port = item.rpartition(":")
        if sep and port.isdigit() and host and "]" not in port:
            pairs.append((host, int(port)))
        else:
            pairs.append((item, args.ssh_port))
    return pairs


def build_commands(args) -> List[List[str]]:
    hosts = host_pairs(args)
    if not hosts:
        raise SystemExit("baguarun: --host_list (or BAGUA_NODE_DOMAIN_NAMES) is required")
    exported = {}
    for item in args.export_env:
        name, sep, value = item.partition("=")
        if sep:
            exported[name] = value
        elif name in os.environ:
            exported[name] = os.environ[name]
    exports = " ".join(f"{k}={shlex.quote(v)}" for k, v in exported.items())
    passthrough = []
    if args.bagua_service_port:
        passthrough.append(f"--bagua_service_port={args.bagua_service_port}")
    if args.no_python:
        passthrough.append("--no_python")
    if args.enable_bagua_net:
        passthrough.append("--enable_bagua_net")
    cmds = []
    for i, (host, port) in enumerate(hosts):
        remote = (f"cd {shlex.quote(os.getcwd())} && {exports} {shlex.quote(sys.executable)} -m bagua_b200.distributed.launch "
                  f"--nnodes={len(hosts)} --node_rank={i} --nproc_per_node={args.nproc_per_node} --master_addr={hosts[0][0]} --master_port={args.master_port} "
                  + " ".join(passthrough + [shlex.quote(a) for a in args.launch_args]))
        cmds.append(["ssh", "-o", "StrictHostKeyChecking=no", "-p", str(port), host, remote])
    return cmds


def main(argv=None) -> int:
    args = parse_args(argv)
    cmds = build_commands(args)
    if args.dry_run:
        for c in cmds:
            print(" ".join(shlex.quote(x) for x in c))
        return 0
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]

    def pump(p, host):
        for line in p.stdout:
            sys.stdout.write(f"[{host}] {line}")

    threads = [threading.Thread(target=pump, args=(p, c[-2]), daemon=True) for p, c in zip(procs, cmds)]
    for t in threads:
        t.start()
    # first failure wins: a host that dies would leave the others blocked in rendezvous / collectives forever, so they are
    # terminated (their launchers forward the signal to the workers)
    rc = 0
    pending = list(procs)
    while pending:
        for p in list(pending):
            r = p.poll()
            if r is None:
                continue
            pending.remove(p)
            if r != 0 and rc == 0:
                rc = r
                for q in pending:
                    q.terminate()
        if pending:
            time.sleep(0.2)
    for t in threads:
        t.join(timeout=2)
    return rc


if __name__ == "__main__":
    sys.exit(main())
