"""Built-in HDF5 support (data/hdf5.py) for the HDF5_DATA / HDF5_OUTPUT layers.

There is no libhdf5 / h5py in the image.  The reader is checked against the one file on the box that a genuine HDF5
library wrote — scipy's MATLAB v7.3 sample (512-byte user block, superblock 0, symbol-table group, v1 object header,
version-2 contiguous layout, float64) — and the writer against the reader."""
import os

import numpy as np
import pytest
import torch

from poseidon_b200 import proto as P
from poseidon_b200.data import hdf5
from poseidon_b200 import Net
from poseidon_b200.proto import parse_text


def _scipy_sample():
    try:
        import scipy.io
    except ImportError:
        return None
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.exists(p) else None


def test_reads_a_file_written_by_the_hdf5_library():
    p = _scipy_sample()
    if p is None:
        pytest.skip("scipy's HDF5 sample file is not installed")
    with hdf5.File(p) as f:
        assert f.keys() == ["testdouble"]
        ds = f.datasets["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8") and ds.layout[0] == "contiguous"
        x = f["testdouble"]
    # the variable scipy's MATLAB test-suite calls `theta`: 0, pi/4, ..., 2*pi
    assert np.allclose(x.reshape(-1), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_roundtrip_contiguous_all_dtypes(tmp_path):
    rng = np.random.RandomState(0)
    arrs = {"data": rng.randn(10, 3, 4, 5).astype(np.float32), "label": np.arange(10, dtype=np.float64).reshape(10, 1),
            "i32": np.arange(-3, 9, dtype=np.int32).reshape(3, 4), "u8": np.arange(7, dtype=np.uint8),
            "be": np.arange(5, dtype=">f4"), "half": np.linspace(0, 1, 6).astype(np.float16)}
    hdf5.save(str(tmp_path / "t.h5"), arrs)
    with hdf5.File(str(tmp_path / "t.h5")) as f:
        assert sorted(f.keys()) == sorted(arrs)
        for k, v in arrs.items():
            got = f[k]
            want = v.astype(np.float32) if v.dtype == np.float16 else v.astype(v.dtype.newbyteorder("="))
            assert got.shape == v.shape and got.dtype == want.dtype and np.array_equal(got, want), k
        with pytest.raises(KeyError):
            f["nope"]
    raw = open(tmp_path / "t.h5", "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0                      # superblock version 0
    assert int.from_bytes(raw[40:48], "little") == len(raw)                      # end-of-file address


@pytest.mark.parametrize("kw", [{}, {"gzip": 4}, {"gzip": 6, "shuffle": True}, {"shuffle": True}])
def test_roundtrip_chunked_filters(tmp_path, kw):
    rng = np.random.RandomState(1)
    arrs = {"data": rng.randn(37, 3, 9, 5).astype(np.float32), "label": np.arange(37, dtype=np.float64).reshape(37, 1),
            "big": (rng.randn(300, 70) * 100).astype(np.int16)}     # 43 x 5 chunks: a two-level chunk B-tree
    path = str(tmp_path / "c.h5")
    hdf5.save(path, arrs, chunks={"data": (8, 3, 4, 5), "big": (7, 16)}, **kw)
    with hdf5.File(path) as f:
        assert f.datasets["big"].layout[0] == "chunked" and f.datasets["label"].layout[0] == "contiguous"
        assert [fid for fid, _ in f.datasets["data"].filters] == ([2] if kw.get("shuffle") else []) + ([1] if kw.get("gzip") else [])
        for k, v in arrs.items():
            assert np.array_equal(f[k], v), k
    if kw.get("gzip"):
        plain = str(tmp_path / "p.h5")
        hdf5.save(plain, {"z": np.zeros((64, 64), np.float32)}, chunks={"z": (16, 64)})
        hdf5.save(path, {"z": np.zeros((64, 64), np.float32)}, chunks={"z": (16, 64)}, **kw)
        assert os.path.getsize(path) < os.path.getsize(plain) - 10000


def test_rejects_what_it_does_not_implement(tmp_path):
    (tmp_path / "x.h5").write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(IOError, match="not an HDF5 file"):
        hdf5.File(str(tmp_path / "x.h5"))
    hdf5.save(str(tmp_path / "v.h5"), {"data": np.zeros(3, np.float32)})
    raw = bytearray(open(tmp_path / "v.h5", "rb").read())
    raw[8] = 2                                                                   # pretend libver='latest'
    (tmp_path / "v2.h5").write_bytes(raw)
    with pytest.raises(IOError, match="superblock version 2"):
        hdf5.File(str(tmp_path / "v2.h5"))
    with pytest.raises(IOError):
        hdf5.save(str(tmp_path / "s.h5"), {"s": np.array(["a", "b"])} | {str(i): np.zeros(1) for i in range(40)})


def test_hdf5_data_and_output_layers_with_real_h5_files(tmp_path):
    """HDF5_DATA walks the files of the source list row by row (reference: hdf5_data_layer.cpp:75-105); HDF5_OUTPUT's
    file is readable back as an HDF5_DATA source."""
    a = {"data": np.arange(24, dtype=np.float32).reshape(6, 1, 2, 2), "label": np.arange(6, dtype=np.float32)}
    b = {"data": 100 + np.arange(16, dtype=np.float64).reshape(4, 1, 2, 2), "label": 10 + np.arange(4, dtype=np.int32)}
    hdf5.save(str(tmp_path / "a.h5"), a)
    hdf5.save(str(tmp_path / "b.h5"), b, chunks={"data": (3, 1, 2, 2)}, gzip=4, shuffle=True)
    (tmp_path / "list.txt").write_text(f"{tmp_path / 'a.h5'}\n{tmp_path / 'b.h5'}\n")
    out_file = tmp_path / "out.h5"
    txt = f'''layers {{ name: "h" type: HDF5_DATA top: "data" top: "label"
                       hdf5_data_param {{ source: "{tmp_path / "list.txt"}" batch_size: 4 }} }}
              layers {{ name: "o" type: HDF5_OUTPUT bottom: "data" bottom: "label"
                       hdf5_output_param {{ file_name: "{out_file}" }} }}'''
    net = Net(parse_text(txt, P.NetParameter), phase=P.TRAIN)
    seen = []
    for _ in range(3):
        net.forward()
        seen += net.blobs["label"].reshape(-1).tolist()
    assert seen == [0, 1, 2, 3, 4, 5, 10, 11, 12, 13, 0, 1]
    assert net.blobs["data"].shape == (4, 1, 2, 2) and net.blobs["data"].dtype == torch.float32
    with hdf5.File(str(out_file)) as f:
        assert f["data"].shape == (12, 1, 2, 2) and f["label"].reshape(-1).tolist() == seen
        assert f["data"][6, 0, 0, 0] == 100.0
    # a file without the expected datasets is named in the error
    hdf5.save(str(tmp_path / "bad.h5"), {"x": np.zeros(2, np.float32)})
    (tmp_path / "bad.txt").write_text(f"{tmp_path / 'bad.h5'}\n")
    bad = f'layers {{ name: "h" type: HDF5_DATA top: "data" top: "label" hdf5_data_param {{ source: "{tmp_path / "bad.txt"}" batch_size: 1 }} }}'
    with pytest.raises(IOError, match="dataset 'data' not found"):
        Net(parse_text(bad, P.NetParameter), phase=P.TRAIN)


def test_corrupt_files_raise_one_exception_type(tmp_path):
    """Mutated files end in HDF5Error (an IOError) or read fine — no stray IndexError / zlib.error / MemoryError."""
    from test_host_fuzz import _mutations
    rng = np.random.RandomState(0)
    hdf5.save(str(tmp_path / "a.h5"), {"data": rng.randn(20, 3, 4, 4).astype(np.float32), "label": np.arange(20.0)},
              chunks={"data": (6, 3, 4, 4)}, gzip=4, shuffle=True)
    raw = (tmp_path / "a.h5").read_bytes()
    outcomes = {"ok": 0, "error": 0}
    for mut in _mutations(raw, rng, 400):
        (tmp_path / "m.h5").write_bytes(mut)
        try:
            with hdf5.File(str(tmp_path / "m.h5")) as f:
                for k in f.keys():
                    f[k]
            outcomes["ok"] += 1
        except hdf5.HDF5Error:
            outcomes["error"] += 1
    assert outcomes["error"] > 100 and outcomes["ok"] > 10, outcomes
