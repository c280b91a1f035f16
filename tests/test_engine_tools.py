"""CaffeEngine facade + dataset converters."""
import gzip
import os
import struct

import numpy as np
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
