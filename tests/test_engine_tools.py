"""CaffeEngine facade + dataset converters."""
import gzip
import json
import os
import struct

import numpy as np
import pytest
import torch

from poseidon_b200 import CaffeEngine, proto as P
from poseidon_b200.data.db import RecordReader
from poseidon_b200.models import zoo
from poseidon_b200.parallel.context import RankContext
from poseidon_b200.tools import convert_cifar_data, convert_mnist_data


def test_convert_mnist(tmp_path):
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, (7, 28, 28)).astype(np.uint8)
    labs = rng.randint(0, 10, 7).astype(np.uint8)
    ip, lp = tmp_path / "img-idx3-ubyte.gz", tmp_path / "lab-idx1-ubyte"
    with gzip.open(ip, "wb") as f:
        f.write(struct.pack(">IIII", 2051, 7, 28, 28) + imgs.tobytes())
    lp.write_bytes(struct.pack(">II", 2049, 7) + labs.tobytes())
    assert convert_mnist_data.main([str(ip), str(lp), str(tmp_path / "mnist_db")]) == 0
    r = RecordReader(str(tmp_path / "mnist_db" / "data.pdb"))
    assert len(r) == 7 and r.key(3) == b"00000003"
    d = r.datum(3)
    assert (d.channels, d.height, d.width, d.label) == (1, 28, 28, int(labs[3]))
    assert np.array_equal(np.frombuffer(d.data, np.uint8).reshape(28, 28), imgs[3])


def test_convert_cifar(tmp_path):
    rng = np.random.RandomState(1)
    src = tmp_path / "bin"
    os.makedirs(src)
    recs = {}
    for fn in [f"data_batch_{i}.bin" for i in range(1, 6)] + ["test_batch.bin"]:
        raw = rng.randint(0, 256, (3, 3073)).astype(np.uint8)
        raw[:, 0] %= 10
        raw.tofile(src / fn)
        recs[fn] = raw
    assert convert_cifar_data.main([str(src), str(tmp_path / "out")]) == 0
    tr = RecordReader(str(tmp_path / "out" / "cifar10_train_db" / "data.pdb"))
    te = RecordReader(str(tmp_path / "out" / "cifar10_test_db" / "data.pdb"))
    assert len(tr) == 15 and len(te) == 3
    d = tr.datum(4)                       # second record of data_batch_2
    assert d.label == int(recs["data_batch_2.bin"][1, 0])
    assert np.array_equal(np.frombuffer(d.data, np.uint8), recs["data_batch_2.bin"][1, 1:])


def test_caffe_engine_plan_and_start(tmp_path):
    net = zoo.lenet(batch=4, test_batch=4)
    sp = zoo.get_solver_param("lenet", net=net, max_iter=2, display=0, snapshot=0, snapshot_after_train=False,
                              test_interval=0, random_seed=3, solver_mode="CPU")
    sp.clear("test_iter")
    eng = CaffeEngine(sp, rank_ctx=RankContext(device=torch.device("cpu")), engine="torch")
    plan = eng.plan()
    names = [(t.layer, t.blob) for t in plan]
    assert ("conv1", 0) in names and ("ip2", 1) in names
    assert [t.global_id for t in plan] == list(range(len(plan)))
    assert all(t.route == "local" for t in plan)
    assert sum(t.count for t in plan) == sum(p.numel() for p in eng.solver.net.params)
    solver = eng.start(net_outputs=str(tmp_path / "run"))
    assert solver.iter == 2
    assert os.path.exists(str(tmp_path / "run") + ".netoutputs")
    eng.close()


def test_zoo_cli_writes_prototxts_that_train(tmp_path):
    """`python -m poseidon_b200.models.zoo` (used by examples/*.sh) + caffe_main train on the written solver."""
    from poseidon_b200.models import zoo
    from poseidon_b200.tools import caffe_main
    out = tmp_path / "models"
    assert zoo.main(["--out", str(out), "--only", "lenet,cifar10_quick"]) == 0
    assert sorted(os.listdir(out)) == ["cifar10_quick", "lenet"]
    sp = P.read_solver(str(out / "cifar10_quick" / "solver.prototxt"))
    assert abs(sp.base_lr - 0.0007) < 1e-9 and sp.max_iter == 4000 and list(sp.test_iter) == [100]
    # shorten the LeNet solver and run the CLI on CPU for two iterations
    lp = out / "lenet" / "solver.prototxt"
    sp = P.read_solver(str(lp))
    sp.max_iter, sp.display, sp.snapshot, sp.test_interval = 2, 1, 0, 0
    sp.snapshot_after_train = False
    sp.clear("test_iter")
    sp.solver_mode = "CPU"
    P.write_text(str(lp), sp)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert caffe_main.main(["train", f"--solver={lp}", f"--net_outputs={tmp_path / 'run'}"]) == 0
    finally:
        os.chdir(cwd)
    assert os.path.exists(str(tmp_path / "run") + ".netoutputs")


def _free_ports(n):
    import socket
    socks = [socket.socket() for _ in range(n)]
    for s in socks:
        s.bind(("127.0.0.1", 0))
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def _lenet_job(tmp_path, max_iter):
    from poseidon_b200 import proto as P
    from poseidon_b200.models import zoo
    net_path = tmp_path / "lenet.prototxt"
    P.write_text(str(net_path), zoo.lenet(batch=4, test_batch=4))
    sp = zoo.lenet_solver(net_path=str(net_path), max_iter=max_iter, display=1, test_interval=0, solver_mode="CPU",
                          snapshot=0, snapshot_prefix=str(tmp_path / "lenet"))
    sp.clear("test_iter")
    P.write_text(str(tmp_path / "solver.prototxt"), sp)
    p0, p1 = _free_ports(2)
    (tmp_path / "hosts").write_text(f"0 127.0.0.1 {p0}\n1 127.0.0.1 {p1}\n")
    return str(tmp_path / "solver.prototxt"), str(tmp_path / "hosts")


def test_launcher_runs_one_process_per_hostfile_line(tmp_path, capsys):
    """tools.launch: two hostfile lines on 127.0.0.1 = a 2-process gloo job (the reference's way of simulating a cluster
    on one box), logs per client, exit code of the job; --dry_run prints the ssh fan-out for remote hosts."""
    from poseidon_b200.tools import launch
    solver, hosts = _lenet_job(tmp_path, 3)
    run = str(tmp_path / "run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rc = launch.main(["train", "--hostfile", hosts, "--solver", solver, "--run_dir", run, "--workdir", root,
                      "--env", "OMP_NUM_THREADS=2", "--", "--svb=true", "--comm=gloo"])
    logs = [open(os.path.join(run, f"client_{i}.log")).read() for i in range(2)]
    assert rc == 0, logs
    assert "Optimization Done" in logs[0] and "world_size=2" in logs[0]
    assert os.path.exists(tmp_path / "lenet_iter_3.caffemodel")
    recs = json.load(open(os.path.join(run, "pids.json")))["clients"]
    assert [r["client"] for r in recs] == [0, 1] and all(r["local"] for r in recs)
    capsys.readouterr()
    assert launch.main(["train", "--nproc_per_node", "3", "--solver", solver, "--dry_run", "--run_dir", run]) == 0
    out3 = capsys.readouterr().out
    assert out3.count("--client_id=") == 3 and len(open(os.path.join(run, "hostfile")).read().splitlines()) == 3
    (tmp_path / "remote").write_text("0 10.0.0.1 9999\n1 10.0.0.2 9999\n")
    assert launch.main(["train", "--hostfile", str(tmp_path / "remote"), "--solver", solver, "--dry_run",
                        "--run_dir", run, "--", "--table_staleness=1"]) == 0
    out = capsys.readouterr().out
    assert out.count("ssh ") == 2 and "10.0.0.2" in out and "--client_id=1" in out and "--table_staleness=1" in out


def test_launcher_fails_fast_and_kill_stops_recorded_processes(tmp_path):
    import subprocess
    import sys
    import time
    from poseidon_b200.tools import launch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # a client that cannot start (missing solver) takes the whole job down with a non-zero exit code
    _, hosts = _lenet_job(tmp_path, 3)
    rc = launch.main(["train", "--hostfile", hosts, "--solver", str(tmp_path / "missing.prototxt"),
                      "--run_dir", str(tmp_path / "bad"), "--workdir", root])
    assert rc != 0
    # a long job is stopped by `kill` through the recorded PIDs
    solver, hosts = _lenet_job(tmp_path, 1000000)
    run = str(tmp_path / "long")
    sup = subprocess.Popen([sys.executable, "-m", "poseidon_b200.tools.launch", "train", "--hostfile", hosts, "--solver",
                            solver, "--run_dir", run, "--workdir", root, "--", "--comm=gloo"], cwd=root,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        deadline = time.time() + 120
        while time.time() < deadline:
            log0 = os.path.join(run, "client_0.log")
            if os.path.exists(log0) and "Iteration" in open(log0).read():
                break
            time.sleep(0.5)
        else:
            raise AssertionError("job did not start")
        assert launch.main(["kill", "--run_dir", run]) == 0
        out, _ = sup.communicate(timeout=60)
        assert sup.returncode != 0 and "stopping the others" in out or "exit code" in out
    finally:
        if sup.poll() is None:
            sup.kill()
    recs = json.load(open(os.path.join(run, "pids.json")))["clients"]
    time.sleep(0.5)
    for r in recs:
        with pytest.raises(ProcessLookupError):
            os.kill(r["pid"], 0)


def test_launcher_restarts_from_latest_snapshot(tmp_path):
    """Rank 1 is killed before iteration 5 of the first attempt; the supervisor stops rank 0, finds lenet_iter_4
    .solverstate and relaunches both; the job then runs to max_iter."""
    from poseidon_b200 import proto as P
    from poseidon_b200.tools import launch
    solver, hosts = _lenet_job(tmp_path, 8)
    sp = P.read_solver(solver)
    sp.snapshot = 2
    P.write_text(solver, sp)
    run = str(tmp_path / "run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rc = launch.main(["train", "--hostfile", hosts, "--solver", solver, "--run_dir", run, "--workdir", root,
                      "--max_restarts", "2", "--env", "POSEIDON_FAULT=kill:rank=1,step=5,attempt=0",
                      "--env", "OMP_NUM_THREADS=2", "--", "--comm=gloo"])
    first = open(os.path.join(run, "client_0.log")).read()
    second = open(os.path.join(run, "client_0.restart1.log")).read()
    assert rc == 0, (first[-1500:], second[-1500:])
    assert "Optimization Done" not in first and "Optimization Done" in second
    assert "lenet_iter_4.solverstate" in second                     # resumed from the newest snapshot
    assert os.path.exists(tmp_path / "lenet_iter_8.caffemodel")
    assert launch.latest_solverstate(solver).endswith("lenet_iter_8.solverstate")


def test_launcher_fused_backend_lenet_with_restart_on_emulation(tmp_path):
    """Everything at once, on the CPU: tools.launch starts 2 ranks of caffe_main with the sm100 engine (kernels
    emulated), the fused backend on shared-memory peer arenas, LeNet (channel-padded layers), SFB on, test phases and
    snapshots; rank 1 is killed before iteration 5; the supervisor restarts both from lenet_iter_4.solverstate."""
    from poseidon_b200 import proto as P
    from poseidon_b200.tools import launch
    solver, hosts = _lenet_job(tmp_path, 8)
    sp = P.read_solver(solver)
    sp.snapshot, sp.test_interval = 2, 4
    sp.test_iter = [2]
    P.write_text(solver, sp)
    run = str(tmp_path / "run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rc = launch.main(["train", "--hostfile", hosts, "--solver", solver, "--run_dir", run, "--workdir", root,
                      "--max_restarts", "1", "--env", "POSEIDON_EMULATE=1", "--env", "OMP_NUM_THREADS=2",
                      "--env", "POSEIDON_FAULT=kill:rank=1,step=5,attempt=0",
                      "--", "--engine=sm100", "--comm=fused", "--svb=true"])
    first = open(os.path.join(run, "client_0.log")).read()
    second = open(os.path.join(run, "client_0.restart1.log")).read()
    assert rc == 0, (first[-1500:], second[-1500:])
    assert "engine=sm100, comm=fused" in first and "Test net output #0: accuracy" in first
    assert "Restored solver state from" in second and "lenet_iter_4.solverstate" in second
    assert "Optimization Done" in second and os.path.exists(tmp_path / "lenet_iter_8.caffemodel")
