"""Native (C++) record loader vs the pure-Python DBSource: same cursor semantics, same bytes."""
import os

import numpy as np
import pytest
import torch

from poseidon_b200 import proto as P
from poseidon_b200.data import native
from poseidon_b200.data.db import RecordReader, RecordWriter
from poseidon_b200.data.source import DBSource

pytestmark = pytest.mark.skipif(not native.available(), reason="host extension could not be built")


def _write_db(path, n, shape, floats=False, seed=0):
    rng = np.random.RandomState(seed)
    c, h, w = shape
    with RecordWriter(path) as wr:
        for i in range(n):
            d = P.Datum()
            d.channels, d.height, d.width = c, h, w
            d.label = int(rng.randint(0, 10)) if i % 7 else -1
            if floats:
                d.float_data = rng.randn(c * h * w).astype(np.float32)
            else:
                d.data = rng.randint(0, 256, c * h * w).astype(np.uint8).tobytes()
            wr.put(f"{i:08d}", d.SerializeToString())


@pytest.mark.parametrize("floats", [False, True])
@pytest.mark.parametrize("offset,stride", [(0, 1), (1, 3)])
def test_native_matches_python(tmp_path, floats, offset, stride):
    path = str(tmp_path / "data.pdb")
    _write_db(path, 23, (3, 5, 4), floats)
    py_src = DBSource(RecordReader(path), 8, offset, stride)
    nat = native.NativeDBSource(path, 8, offset, stride, threads=3, depth=3, pin=False)
    assert nat.shape == (3, 5, 4) and nat.is_bytes == (not floats) and len(nat) == 23
    for _ in range(9):                 # > 3 epochs: wrap-around and ring reuse
        xa, ya = py_src.next_batch()
        xb, yb = nat.next_batch()
        assert xa.dtype == xb.dtype
        assert torch.equal(xa, xb)
        assert torch.equal(ya, yb)
    nat.close()


def test_sync_fill_and_seek(tmp_path):
    path = str(tmp_path / "data.pdb")
    _write_db(path, 10, (1, 2, 2))
    m = native.module()
    ld = m.BatchLoader(path, 4, 2, 1, 2)
    x = torch.empty(4, 1, 2, 2, dtype=torch.uint8)
    y = torch.empty(4)
    ld.fill(x.data_ptr(), y.data_ptr())
    assert ld.position() == 6
    ref = DBSource(RecordReader(path), 4, 2, 1).next_batch()
    assert torch.equal(x, ref[0]) and torch.equal(y, ref[1])
    ld.seek(8)
    ld.fill(x.data_ptr(), y.data_ptr())
    assert ld.position() == 2      # wrapped


def test_bad_file(tmp_path):
    p = tmp_path / "junk.pdb"
    p.write_bytes(b"not a database at all")
    with pytest.raises(RuntimeError):
        native.module().BatchLoader(str(p), 2, 0, 1, 1)


def test_bf16_wire_roundtrip():
    x = torch.randn(1000) * 100
    x[0], x[1], x[2] = float("inf"), float("nan"), 0.0
    b = native.f32_to_bf16(x)
    ref = x.to(torch.bfloat16)
    assert torch.equal(b[2:].view(torch.int16), ref[2:].view(torch.int16))
    assert torch.isinf(b[0]) and torch.isnan(b[1])
    assert torch.equal(native.bf16_to_f32(b)[2:], ref[2:].float())


def test_data_layer_uses_native_loader(tmp_path):
    from poseidon_b200.layers.base import NetContext
    from poseidon_b200.net.net import Net
    db = tmp_path / "train_db"
    os.makedirs(db)
    _write_db(str(db / "data.pdb"), 12, (3, 8, 8))
    net_txt = f'''
    name: "t"
    layers {{ name: "data" type: DATA top: "data" top: "label"
             data_param {{ source: "{db}" batch_size: 4 }} transform_param {{ crop_size: 6 }} }}
    '''
    f = tmp_path / "net.prototxt"
    f.write_text(net_txt)
    npar = P.read_net(str(f))
    net = Net(npar, phase=P.TRAIN, ctx=NetContext())
    out = net.forward()
    dl = net.data_layers()[0]
    assert type(dl.source).__name__ == "NativeDBSource"
    assert net.blobs["data"].shape == (4, 3, 6, 6)
    net.close()
