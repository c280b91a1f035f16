"""Cluster launcher: one ``caffe_main`` process per hostfile line, started locally or over ssh, logs and PIDs under one
run directory, fail-fast supervision, and a ``kill`` command that stops exactly the processes it started.

reference: examples/{mnist,cifar10,imagenet,googlenet}/run_local.py (builds the gflags command line for one client and
``os.system``s it; the operator runs it on every machine), examples/imagenet/train_imagenet.sh:53-105 (ssh fan-out over
the hostfile), scripts/kill_caffe.py (``killall caffe_main`` on every host), machinefiles/localserver (``0 127.0.0.1
9999``).  Differences on purpose:

* the hostfile has one line per *process* = per GPU (``<id> <ip> <port>``; repeat a host for each of its GPUs, as the
  reference does to simulate several nodes on one box); line 0 is the rendezvous;
* a client that dies takes the job down (the reference's peers would wait forever in the PS clock);
* ``kill`` signals the recorded PIDs (and their process groups) instead of every process with a matching name;
* ``--max_restarts N``: a failed job is relaunched from its newest ``.solverstate`` (the reference's manual recovery).

    python -m poseidon_b200.tools.launch train --hostfile machinefiles/localserver --solver models/lenet/solver.prototxt \\
        --run_dir output/lenet -- --svb=true --table_staleness=0
    python -m poseidon_b200.tools.launch kill --run_dir output/lenet
"""
from __future__ import annotations

import argparse
import json
import os
import shlex
import signal
import socket
import subprocess
import sys
import time
from typing import Dict, List

from ..parallel.context import parse_hostfile

_LOCAL = {"127.0.0.1", "localhost", "::1"}


def _is_local(ip: str) -> bool:
    if ip in _LOCAL:
        return True
    try:
        return ip in (socket.gethostname(), socket.gethostbyname(socket.gethostname()))
    except OSError:
        return False


def client_command(args, client_id: int, extra: List[str]) -> List[str]:
    cmd = [args.python, "-m", "poseidon_b200.tools.caffe_main", args.command, f"--hostfile={args.hostfile}",
           f"--client_id={client_id}"]
    if args.solver:
        cmd.append(f"--solver={args.solver}")
    if args.snapshot:
        cmd.append(f"--snapshot={args.snapshot}")
    if args.weights:
        cmd.append(f"--weights={args.weights}")
    if args.net_outputs:
        cmd.append(f"--net_outputs={args.net_outputs}")
    return cmd + list(extra)


def _remote_shell(args, ip: str, cmd: List[str], pidfile: str, env: Dict[str, str]) -> List[str]:
    exports = " ".join(f"{k}={shlex.quote(v)}" for k, v in env.items())
    inner = f"cd {shlex.quote(args.workdir)} && echo $$ > {shlex.quote(pidfile)} && exec env {exports} " + \
        " ".join(shlex.quote(c) for c in cmd)
    return shlex.split(args.ssh) + [ip, f"sh -c {shlex.quote(inner)}"]


def latest_solverstate(solver_path: str):
    """Newest ``<snapshot_prefix>_iter_N.solverstate`` of the solver's run (None if there is none yet)."""
    import glob
    import re
    from .. import proto as P
    from ..utils.paths import expand_placeholder
    sp = P.read_solver(solver_path)
    prefix = expand_placeholder(sp.snapshot_prefix or "snapshot", os.path.dirname(os.path.abspath(solver_path)),
                                must_exist=False)
    best = (-1, None)
    # synchronous modes write one `<prefix>_iter_N.solverstate`; the bounded-staleness modes write per-worker files
    # `<prefix>_iter_N.solverstate.<rank>.0` (momentum is per worker there) — Solver.restore takes the UNSUFFIXED base name
    # and picks its own rank's file, so that is what is returned in both cases
    for f in glob.glob(glob.escape(prefix) + "_iter_*.solverstate*"):
        m = re.search(r"^(.*_iter_(\d+)\.solverstate)(\.\d+\.\d+)?$", f)
        if m and int(m.group(2)) > best[0]:
            best = (int(m.group(2)), m.group(1))
    return best[1]


def cmd_train(args, extra: List[str]) -> int:
    """Run the job; with --max_restarts, a failed job is restarted from its newest snapshot (the reference's recovery
    procedure — kill everything, relaunch with --snapshot — done by the supervisor instead of the operator)."""
    attempt = 0
    while True:
        rc = _run_once(args, extra, attempt)
        if rc == 0 or rc == 130 or args.dry_run or attempt >= args.max_restarts or args.command != "train":
            return rc
        attempt += 1
        state = latest_solverstate(args.solver) if args.solver else None
        if state:
            args.snapshot, args.weights = state, ""
        print(f"[launch] restart {attempt}/{args.max_restarts} from {state or 'scratch (no snapshot yet)'}", flush=True)


def _run_once(args, extra: List[str], attempt: int) -> int:
    hosts = parse_hostfile(args.hostfile)
    if not hosts:
        raise SystemExit(f"{args.hostfile}: no hosts")
    os.makedirs(args.run_dir, exist_ok=True)
    env_extra = dict(kv.split("=", 1) for kv in args.env)
    env_extra["POSEIDON_ATTEMPT"] = str(attempt)
    procs, records = [], []
    for cid, ip, port in hosts:
        cmd = client_command(args, cid, extra)
        log_path = os.path.join(args.run_dir, f"client_{cid}.log" if attempt == 0 else f"client_{cid}.restart{attempt}.log")
        local = _is_local(ip) and not args.force_ssh
        pidfile = os.path.join(args.run_dir, f"client_{cid}.pid")
        full = cmd if local else _remote_shell(args, ip, cmd, pidfile, env_extra)
        print(f"[launch] client {cid} on {ip}:{port}: {' '.join(shlex.quote(c) for c in full)}", flush=True)
        if args.dry_run:
            continue
        log = open(log_path, "w")
        p = subprocess.Popen(full, stdout=log, stderr=subprocess.STDOUT, cwd=args.workdir if local else None,
                             env=dict(os.environ, **env_extra) if local else None, start_new_session=True)
        procs.append((cid, p, log))
        records.append({"client": cid, "ip": ip, "local": local, "pid": p.pid, "pidfile": None if local else pidfile})
    if args.dry_run:
        return 0
    with open(os.path.join(args.run_dir, "pids.json"), "w") as f:
        json.dump({"ssh": args.ssh, "clients": records}, f, indent=1)
    # supervise: first failure stops everyone
    rc = 0
    alive = {cid: p for cid, p, _ in procs}
    try:
        while alive:
            for cid, p in list(alive.items()):
                r = p.poll()
                if r is None:
                    continue
                del alive[cid]
                if r != 0:
                    rc = rc or r
                    print(f"[launch] client {cid} exited with {r}; stopping the others (log: "
                          f"{os.path.join(args.run_dir, f'client_{cid}.log')})", flush=True)
                    _stop(records, args.ssh, only=set(alive))
            time.sleep(0.2)
    except KeyboardInterrupt:
        _stop(records, args.ssh)
        rc = 130
    for _, _, log in procs:
        log.close()
    print(f"[launch] job finished with exit code {rc}", flush=True)
    return rc


def _stop(records, ssh: str, only=None, sig=signal.SIGTERM) -> int:
    n = 0
    for r in records:
        if only is not None and r["client"] not in only:
            continue
        try:
            if r["local"]:
                os.killpg(r["pid"], sig)                      # start_new_session: pid == process-group id
            else:
                pf = shlex.quote(r["pidfile"])
                remote = f"kill -{int(sig)} -- -$(cat {pf}) 2>/dev/null || kill -{int(sig)} $(cat {pf})"
                subprocess.call(shlex.split(ssh) + [r["ip"], remote])
                os.kill(r["pid"], sig)                         # the local ssh client
            n += 1
        except (ProcessLookupError, PermissionError, FileNotFoundError):
            pass
    return n


def cmd_kill(args) -> int:
    path = os.path.join(args.run_dir, "pids.json")
    if not os.path.exists(path):
        raise SystemExit(f"{path}: no record of a launched job")
    with open(path) as f:
        rec = json.load(f)
    n = _stop(rec["clients"], rec.get("ssh", "ssh"), sig=signal.SIGKILL if args.force else signal.SIGTERM)
    print(f"[launch] signalled {n} of {len(rec['clients'])} clients")
    return 0


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    extra: List[str] = []
    if "--" in argv:
        i = argv.index("--")
        argv, extra = argv[:i], argv[i + 1:]
    ap = argparse.ArgumentParser(prog="poseidon_b200.tools.launch", description=__doc__.split("\n\n")[0])
    ap.add_argument("command", choices=["train", "test", "kill"])
    ap.add_argument("--hostfile", default="machinefiles/localserver")
    ap.add_argument("--run_dir", default="output/run", help="logs (client_<id>.log) and pids.json")
    ap.add_argument("--solver", default="")
    ap.add_argument("--snapshot", default="")
    ap.add_argument("--weights", default="")
    ap.add_argument("--net_outputs", default="")
    ap.add_argument("--python", default=sys.executable)
    ap.add_argument("--workdir", default=os.getcwd(), help="directory to run in on every host (shared checkout)")
    ap.add_argument("--ssh", default="ssh -o StrictHostKeyChecking=no -o BatchMode=yes")
    ap.add_argument("--force_ssh", action="store_true", help="use ssh for local addresses too")
    ap.add_argument("--env", action="append", default=[], metavar="K=V", help="environment for every client (repeatable)")
    ap.add_argument("--dry_run", action="store_true", help="print the per-host command lines and exit")
    ap.add_argument("--max_restarts", type=int, default=0,
                    help="restart a failed training job up to N times from its newest .solverstate")
    ap.add_argument("--force", action="store_true", help="kill: SIGKILL instead of SIGTERM")
    ap.add_argument("--nproc_per_node", type=int, default=0,
                    help="no hostfile needed: run N processes on this machine (writes run_dir/hostfile)")
    args = ap.parse_args(argv)
    if args.command == "kill":
        return cmd_kill(args)
    if args.nproc_per_node > 0:
        os.makedirs(args.run_dir, exist_ok=True)
        args.hostfile = os.path.join(args.run_dir, "hostfile")
        s0 = socket.socket()
        s0.bind(("127.0.0.1", 0))
        port = s0.getsockname()[1]
        s0.close()
        with open(args.hostfile, "w") as f:
            f.write("".join(f"{i} 127.0.0.1 {port + i}\n" for i in range(args.nproc_per_node)))
    return cmd_train(args, extra)


if __name__ == "__main__":
    sys.exit(main())
