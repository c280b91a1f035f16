"""poseidon_b200.ml — the generic ML-application helpers of the PMLS tree (reference: ps/src/ml/)."""
import math

import numpy as np
import pytest
import torch

from poseidon_b200 import ml


def _write_libsvm(path, n, dim, seed=0, one_based=False):
    rng = np.random.RandomState(seed)
    dense = np.zeros((n, dim), np.float32)
    labels = rng.randint(0, 2, size=n)
    with open(path, "w") as f:
        for i in range(n):
            ids = np.sort(rng.choice(dim, rng.randint(0, 6), replace=False))
            vals = rng.randn(len(ids)).astype(np.float32)
            dense[i, ids] = vals
            f.write(str(labels[i] + (1 if one_based else 0)) + "".join(f" {a + (1 if one_based else 0)}:{float(v)!r}" for a, v in zip(ids, vals)) + "\n")
            if i % 7 == 0:
                f.write("# a comment line\n\n")
    return dense, labels


def test_libsvm_loader_matches_dense_ground_truth(tmp_path):
    dense, labels = _write_libsvm(tmp_path / "a.svm", 500, 40, one_based=True)
    batch, lab = ml.read_data_label_libsvm(str(tmp_path / "a.svm"), 40, feature_one_based=True, label_one_based=True)
    assert len(batch) == 500 and lab.tolist() == labels.tolist()
    assert np.array_equal(batch.to_dense().numpy(), dense)
    assert torch.equal(batch.to_sparse_csr().to_dense(), batch.to_dense())
    w = torch.randn(40)
    assert torch.allclose(batch.matvec(w), torch.from_numpy(dense) @ w, atol=1e-5)
    # first 10 samples only; dense output; single parser thread gives the same arrays
    few, lab10 = ml.read_data_label_libsvm(str(tmp_path / "a.svm"), 40, 10, True, True, dense=True, threads=1)
    assert few.shape == (10, 40) and np.array_equal(few.numpy(), dense[:10]) and lab10.tolist() == labels[:10].tolist()
    row = batch[3]
    assert isinstance(row, ml.SparseFeature) and np.array_equal(row.to_dense().numpy(), dense[3])
    (tmp_path / "bad.svm").write_text("1 3:0.5 oops\n")
    with pytest.raises(RuntimeError, match="libsvm"):
        ml.read_data_label_libsvm(str(tmp_path / "bad.svm"), 10)
    with pytest.raises(ValueError, match="feature_dim"):
        ml.read_data_label_libsvm(str(tmp_path / "a.svm"), 5, feature_one_based=True, label_one_based=True)


def test_binary_formats_roundtrip(tmp_path):
    rng = np.random.RandomState(1)
    x = rng.randn(12, 7).astype(np.float32)
    y = rng.randint(1, 4, size=12).astype(np.int32)
    with open(tmp_path / "d.bin", "wb") as f:
        for i in range(12):
            f.write(y[i].tobytes() + x[i].tobytes())
    feats, labels = ml.read_data_label_binary(str(tmp_path / "d.bin"), 7, 12, label_one_based=True)
    assert np.array_equal(feats.numpy(), x) and labels.tolist() == (y - 1).tolist()
    with pytest.raises(IOError):
        ml.read_data_label_binary(str(tmp_path / "d.bin"), 7, 13)
    dense, lab = _write_libsvm(tmp_path / "s.svm", 50, 30, seed=2)
    batch, l1 = ml.read_data_label_libsvm(str(tmp_path / "s.svm"), 30)
    ml.write_sparse_feature_binary(str(tmp_path / "s.bin"), batch, l1)
    b2, l2 = ml.read_data_label_sparse_feature_binary(str(tmp_path / "s.bin"), 30)
    assert np.array_equal(b2.to_dense().numpy(), dense) and l2.tolist() == lab.tolist()
    b3, l3 = ml.read_data_label_sparse_feature_binary(str(tmp_path / "s.bin"), 30, num_data=5)
    assert len(b3) == 5 and l3.tolist() == lab[:5].tolist()


def test_features_and_math():
    d = ml.DenseFeature([1.0, 0.0, 2.0, -1.0])
    s = ml.SparseFeature([3, 0], [4.0, 0.5], 4)          # given unsorted
    assert [i for i, _ in s.entries()] == [0, 3] and s[3] == 4.0 and s[1] == 0.0 and s.num_entries == 2
    s[2] = 7.0
    assert s.num_entries == 3 and s.to_dense().tolist() == [0.5, 0.0, 7.0, 4.0]
    t = ml.SparseFeature([2, 3], [1.0, 1.0], 4)
    assert ml.dot(d, d) == 6.0 and ml.dot(s, d) == ml.dot(d, s) == 0.5 + 14.0 - 4.0 and ml.dot(s, t) == 11.0
    acc = ml.DenseFeature(torch.zeros(4))
    ml.feature_scale_and_add(2.0, s, acc)
    ml.feature_scale_and_add(-1.0, d, acc)
    assert acc.to_dense().tolist() == [0.0, 0.0, 12.0, 9.0]
    with pytest.raises(ValueError):
        ml.SparseFeature([1, 1], [1.0, 2.0], 4)
    with pytest.raises(ValueError):
        ml.SparseFeature([4], [1.0], 4)
    assert ml.safe_log(0.0) == math.log(1e-10) and abs(ml.sigmoid(-800.0)) < 1e-300 and ml.sigmoid(800.0) == 1.0
    assert abs(ml.log_sum(1000.0, 1000.0) - (1000.0 + math.log(2))) < 1e-9
    assert abs(ml.log_sum_vec([0.0, 0.0, 0.0]) - math.log(3)) < 1e-12
    assert torch.allclose(ml.softmax([1.0, 1.0]), torch.tensor([0.5, 0.5]))


def test_workload_manager_partitions_and_wraps():
    # global data: 10 samples over 2 clients x 2 threads -> 3, 3, 2, 2 (disjoint, covering)
    spans = []
    for c in range(2):
        for t in range(2):
            w = ml.WorkloadManager(ml.WorkloadManagerConfig(thread_id=t, client_id=c, num_clients=2, num_threads=2,
                                                            num_batches_per_epoch=2, num_data=10, global_data=True))
            spans.append((w.begin, w.end))
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
    w = ml.WorkloadManager(ml.WorkloadManagerConfig(thread_id=0, client_id=0, num_clients=2, num_threads=2,
                                                    num_batches_per_epoch=2, num_data=10, global_data=True))
    assert w.get_batch_size() == 2 and w.get_num_batches() == 2          # 3 samples, 2 batches of ceil(1.5)
    assert w.get_batch_data_idx(4) == [0, 1, 2, 0]
    seen = []
    while not w.is_end():
        seen.append(w.get_data_idx_and_advance())
        if w.is_end_of_batch():
            seen.append("|")
    assert seen == [0, 1, "|", 2, 0, "|"]
    with pytest.raises(RuntimeError):
        w.get_data_idx_and_advance()
    w.restart()
    assert w.get_data_idx_and_advance() == 0
    # local data: the last thread takes the remainder
    w = ml.WorkloadManager(ml.WorkloadManagerConfig(thread_id=2, num_threads=3, num_batches_per_epoch=1, num_data=10,
                                                    global_data=False))
    assert (w.begin, w.end, w.get_batch_size()) == (6, 10, 4)
    with pytest.raises(ValueError):
        ml.WorkloadManager(ml.WorkloadManagerConfig(thread_id=3, num_threads=4, num_clients=4, client_id=3,
                                                    num_batches_per_epoch=1, num_data=3, global_data=True))


def test_metafile_reader(tmp_path):
    (tmp_path / "data.meta").write_text("num_train_total: 1000\nfeature_dim: 54\nformat: libsvm\n# note\nsparse: true\nscale: 0.5\n")
    m = ml.MetafileReader(str(tmp_path / "data.meta"))
    assert m.get_int32("num_train_total") == 1000 and m.get_string("format") == "libsvm" and m.get_bool("sparse")
    assert m.get_double("scale") == 0.5
    with pytest.raises(KeyError):
        m.get_int32("missing")


def test_disk_streamer_passes_blocks_and_shutdown(tmp_path):
    d0, l0 = _write_libsvm(tmp_path / "p0.svm", 130, 25, seed=3)
    d1, l1 = _write_libsvm(tmp_path / "p1.svm", 70, 25, seed=4)
    dense, labels = np.concatenate([d0, d1]), np.concatenate([l0, l1])
    with ml.DiskStreamer([str(tmp_path / "p0.svm"), str(tmp_path / "p1.svm")], 25, num_passes=2, num_buffers=2,
                         block_bytes=512) as ds:                       # tiny blocks: many hand-offs, line-boundary cuts
        rows, labs = [], []
        while True:
            batch, lab = ds.get_next_data(37)
            if len(batch) == 0:
                break
            assert len(batch) == 37 or len(rows) * 37 + len(batch) == 400
            rows.append(batch.to_dense().numpy())
            labs += lab.tolist()
        got = np.concatenate(rows)
    assert got.shape == (400, 25) and np.array_equal(got, np.concatenate([dense, dense])) and labs == labels.tolist() * 2
    # early shutdown with a full queue does not hang
    ds = ml.DiskStreamer(str(tmp_path / "p0.svm"), 25, num_passes=1000, block_bytes=256)
    b, _ = ds.get_next_data(5)
    assert len(b) == 5
    ds.shutdown()
    assert not ds.thread.is_alive()
