"""Record DB, worker sharding, DataTransformer, data layers and dataset tools."""

import numpy as np
import pytest
import torch

from poseidon_b200 import Net
from poseidon_b200 import proto as P
from poseidon_b200.data.db import RecordWriter, open_db, shard_indices
from poseidon_b200.data.transformer import DataTransformer
from poseidon_b200.proto import parse_text


def _write_db(path, n=10, shape=(3, 8, 8)):
    rng = np.random.RandomState(0)
    with RecordWriter(str(path)) as w:
        for i in range(n):
            img = rng.randint(0, 256, size=shape).astype(np.uint8)
            d = P.Datum(channels=shape[0], height=shape[1], width=shape[2], data=img.tobytes(), label=i % 4)
            w.put("%08d" % i, d.SerializeToString())
    return str(path)


def test_record_db_roundtrip(tmp_path):
    p = _write_db(tmp_path / "db.pdb", 7)
    db = open_db(p)
    assert len(db) == 7 and db.key(3) == b"00000003"
    d = db.datum(5)
    assert (d.channels, d.height, d.width, d.label) == (3, 8, 8, 1) and len(d.data) == 192
    with pytest.raises(IOError):
        open_db(str(tmp_path / "missing"))


def test_shard_rule_matches_reference():
    # shared FS: worker (client*threads+thread) of clients*threads ; private shards: stride by threads only
    assert shard_indices(100, True, 4, 2, 2, 1) == (5, 8)
    assert shard_indices(100, False, 4, 2, 2, 1) == (1, 2)


def test_transformer_center_crop_mean_scale():
    tp = P.TransformationParameter(crop_size=4, scale=0.5)
    tp.mean_value = [10.0]
    tr = DataTransformer(tp, P.TEST)
    x = torch.arange(36, dtype=torch.float32).reshape(1, 1, 6, 6)
    y = tr(x)
    assert y.shape == (1, 1, 4, 4)
    assert torch.allclose(y, (x[:, :, 1:5, 1:5] - 10) * 0.5)
    tp2 = P.TransformationParameter(crop_size=4, mirror=True)
    tr2 = DataTransformer(tp2, P.TRAIN, seed=1)
    outs = {tuple(tr2(x).reshape(-1).tolist()) for _ in range(30)}
    assert len(outs) > 4                                        # random crops + mirrors
    with pytest.raises(ValueError):
        bad = P.TransformationParameter(mean_file="nope.binaryproto")
        bad.mean_value = [1.0]
        DataTransformer(bad, P.TEST)


def test_data_layer_reads_db_with_sharding(tmp_path):
    p = _write_db(tmp_path / "train.pdb", 10)
    txt = f'''layers {{ name: "data" type: DATA top: "data" top: "label"
        data_param {{ source: "{p}" batch_size: 3 shared_file_system: true }} transform_param {{ scale: 0.00390625 }} }}'''
    net = Net(parse_text(txt, P.NetParameter), phase=P.TRAIN)
    _, outs = net.forward()
    assert outs["data"].shape == (3, 3, 8, 8) and outs["label"].reshape(-1).tolist() == [0, 1, 2]
    _, outs = net.forward()
    assert outs["label"].reshape(-1).tolist() == [3, 0, 1]
    net.close()
    from poseidon_b200 import NetContext
    ctx = NetContext(rank=1, world_size=2)
    net2 = Net(parse_text(txt, P.NetParameter), phase=P.TRAIN, ctx=ctx)
    _, outs = net2.forward()
    assert outs["label"].reshape(-1).tolist() == [1, 3, 1]       # records 1, 3, 5
    net2.close()


def test_dummy_memory_and_hdf5_npz_layers(tmp_path):
    txt = '''layers { name: "d" type: DUMMY_DATA top: "a" top: "b" dummy_data_param {
        num: 2 channels: 3 height: 4 width: 4 num: 2 channels: 1 height: 1 width: 1
        data_filler { type: "gaussian" std: 1 } data_filler { type: "constant" value: 3 } } }'''
    net = Net(parse_text(txt, P.NetParameter), phase=P.TRAIN)
    _, o1 = net.forward()
    _, o2 = net.forward()
    assert o1["a"].shape == (2, 3, 4, 4) and not torch.equal(o1["a"], o2["a"]) and torch.all(o2["b"] == 3)
    np.savez(tmp_path / "h.npz", data=np.arange(24, dtype=np.float32).reshape(6, 1, 2, 2), label=np.arange(6, dtype=np.float32))
    (tmp_path / "list.txt").write_text(str(tmp_path / "h.npz") + "\n")
    txt = f'layers {{ name: "h" type: HDF5_DATA top: "data" top: "label" hdf5_data_param {{ source: "{tmp_path / "list.txt"}" batch_size: 4 }} }}'
    net = Net(parse_text(txt, P.NetParameter), phase=P.TRAIN)
    _, o = net.forward()
    assert o["label"].reshape(-1).tolist() == [0, 1, 2, 3]
    _, o = net.forward()
    assert o["label"].reshape(-1).tolist() == [4, 5, 0, 1]


def test_tools_convert_mean_partition(tmp_path):
    import cv2
    from poseidon_b200.tools import compute_image_mean, convert_imageset, partition_data
    rng = np.random.RandomState(1)
    lines = []
    for i in range(6):
        cv2.imwrite(str(tmp_path / f"im{i}.png"), rng.randint(0, 256, size=(10, 12, 3)).astype(np.uint8))
        lines.append(f"im{i}.png {i % 3}")
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    db = str(tmp_path / "imgs.pdb")
    assert convert_imageset.main([str(tmp_path) + "/", str(tmp_path / "list.txt"), db, "--resize_height", "8",
                                  "--resize_width", "8"]) == 0
    r = open_db(db)
    assert len(r) == 6 and r.datum(0).height == 8 and r.datum(4).label == 1
    assert compute_image_mean.main([db, str(tmp_path / "mean.binaryproto")]) == 0
    mean = P.blob_to_array(P.read_binary(str(tmp_path / "mean.binaryproto"), P.BlobProto))
    assert mean.shape == (1, 3, 8, 8) and 60 < mean.mean() < 200
    parts = partition_data.partition(db, 2)
    a, b = open_db(parts[0]), open_db(parts[1])
    assert len(a) == 3 and len(b) == 3 and a.key(1) == r.key(2) and b.key(0) == r.key(1)
    # image-list data layer
    txt = f'''layers {{ name: "d" type: IMAGE_DATA top: "data" top: "label"
        image_data_param {{ source: "{tmp_path / "list.txt"}" batch_size: 2 new_height: 6 new_width: 6 }} }}'''
    net = Net(parse_text(txt, P.NetParameter), phase=P.TEST)
    _, o = net.forward()
    assert o["data"].shape == (2, 3, 6, 6) and o["label"].reshape(-1).tolist() == [0, 1]
    net.close()


def test_zoo_prototxt_export_is_parseable(tmp_path):
    from poseidon_b200.models import zoo
    zoo.write_zoo(str(tmp_path))
    for name in ("alexnet", "googlenet", "lenet", "vgg16", "caffenet"):
        net = P.read_net(str(tmp_path / name / "train_val.prototxt"))
        assert len(net.layers) > 5
        P.read_solver(str(tmp_path / name / "solver.prototxt"))
