"""Bösen-style table API: SSP read guarantees, exactly-once application, BSP at staleness 0 (gloo, 2 ranks)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from test_dist_cpu import _free_port

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("staleness", [0, 2])
def test_table_group_two_workers(tmp_path, staleness):
    port = _free_port()
    out = str(tmp_path / "t")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "ps_worker.py"), out, str(staleness)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=120)
        assert p.returncode == 0, o[-2000:]
    a, b = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert np.allclose(a, b)
    assert a[1, 2] == pytest.approx(5 * 3.0)            # 5 iterations x (1 + 2)
    assert a[2, 0] == pytest.approx(10.0) and a[2, 3] == pytest.approx(5 * 6.0)
    assert np.allclose(a[3], 2 * 15.0)


def test_vector_clock():
    from poseidon_b200.ps import VectorClock
    vc = VectorClock([0, 1, 2])
    assert vc.tick(0) == 0 and vc.tick(1) == 0
    assert vc.tick(2) == 1 and vc.get_min_clock() == 1
    assert vc.tick(2) == 0 and vc.get_clock(2) == 2


def test_api_aliases_single_process():
    import torch
    from poseidon_b200.ps.table import PSTableGroup
    g = PSTableGroup.init(None, staleness=0)
    assert g.register_row(3) == 3 and g.wait_thread_register() is None
    g.turn_on_early_comm(); g.turn_off_early_comm()
    t = g.create_table(0, num_rows=2, row_capacity=6)
    g.create_table_done()
    t.dense_batch_inc(1, torch.tensor([1.0, 2.0, 3.0]), index_st=2)
    t.thread_inc(1, 0, 5.0)
    t.flush_thread_cache()
    g.clock()
    assert t.thread_get(1).tolist() == [5.0, 0.0, 1.0, 2.0, 3.0, 0.0]
    assert t.get_async(1) is None or True
    g.shut_down()


def test_row_types():
    import torch
    from poseidon_b200.ps import DenseRowFloat16, MultiplicativeDenseRow, SortedVectorMapRow, SparseFeatureRow, SparseRow
    r = SparseRow(100)
    r.apply_batch_inc([5, 90, 5], [1.0, 2.0, 0.5])
    r.apply_inc(7, -1.0)
    assert r.copy_to_vector() == [(5, 1.5), (7, -1.0), (90, 2.0)] and r[5].item() == 1.5 and r[6].item() == 0.0
    r.apply_inc(7, 1.0)
    assert r.num_entries() == 3 and r[7].item() == 0.0          # a plain sparse row keeps zero entries (map store)
    s = SortedVectorMapRow(100)
    s.apply_batch_inc([5, 7], [1.0, 1.0])
    s.apply_inc(7, -1.0)
    assert s.copy_to_vector() == [(5, 1.0)]                        # the sorted-vector-map row drops them
    with pytest.raises(IndexError):
        s.apply_inc(100, 1.0)
    f = SparseFeatureRow(10)
    f.apply_dense_batch_inc(torch.tensor([1.0, 2.0]), index_st=3)
    c, v = f.copy_to_tensors()
    assert c.tolist() == [3, 4] and v.tolist() == [1.0, 2.0] and f.to_dense().tolist()[2:6] == [0.0, 1.0, 2.0, 0.0]
    m = MultiplicativeDenseRow(4)
    m.apply_batch_inc([1, 1, 2], [2.0, 3.0, 0.5])
    assert m.to_dense().tolist() == [1.0, 6.0, 0.5, 1.0]
    assert DenseRowFloat16.wire_dtype is torch.float16


def test_adarevision_server_logic_single_worker_matches_the_formulas():
    """One worker, two clocks: the second gradient was computed on the version-1 row, i.e. without knowledge of ... nothing
    (no concurrent updates), so g_bck = 0 and AdaRevision reduces to AdaGrad-style steps; then a delayed gradient (version
    1 applied after version 2 exists) gets the (eta_old - eta) * g_bck correction."""
    import math
    import torch
    from poseidon_b200.ps import AdaRevisionServerTableLogic, PSTableGroup
    g = PSTableGroup.init(None, staleness=0)
    logic = AdaRevisionServerTableLogic(init_step_size=0.5)
    t = g.create_table(0, num_rows=1, row_capacity=2, table_logic=logic)
    g.create_table_done()
    t.batch_inc(0, {0: 2.0})
    assert t.get(0).to_dense().tolist() == [0.0, 0.0]                      # a table with a server logic has no read-my-writes
    g.clock()
    # z = 1 + 2*2 = 5, eta = 0.5 / sqrt(5), delta = -eta * 2
    w1 = -0.5 / math.sqrt(5) * 2.0
    assert t.get(0).to_dense()[0].item() == pytest.approx(w1, rel=1e-6)
    logic.old_accum[(0, 1)][1] += 1                                       # a second client also holds version 1 (the straggler below)
    t.batch_inc(0, {0: 1.0})                                              # computed on version 1: g_bck = accum(2) - old(2) = 0
    g.clock()
    z2 = 5 + 1.0 * (1.0 + 0.0)
    w2 = w1 - 0.5 / math.sqrt(z2) * 1.0
    assert t.get(0).to_dense()[0].item() == pytest.approx(w2, rel=1e-6)
    # a straggler's gradient that was computed on version 1 arrives now: g_bck = accum_now(3) - accum_at_v1(2) = 1
    row = t.rows[0]
    logic.apply_row_oplog(0, torch.tensor([0]), torch.tensor([1.0]), row, row_version=1, end_of_version=True)
    eta_old = 0.5 / math.sqrt(z2)
    z3 = z2 + 1.0 * (1.0 + 2 * 1.0)
    eta = 0.5 / math.sqrt(z3)
    w3 = w2 - eta * 1.0 + (eta_old - eta) * 1.0
    assert row.to_dense()[0].item() == pytest.approx(w3, rel=1e-6)
    g.shut_down()


def test_sparse_and_adarevision_tables_two_workers(tmp_path):
    port = _free_port()
    out = str(tmp_path / "r")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "ps_rows_worker.py"), out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=120)
        assert p.returncode == 0, o[-2000:]
    a, b = dict(np.load(out + ".0.npz")), dict(np.load(out + ".1.npz"))
    for k in a:
        assert np.allclose(a[k], b[k]), k                                    # replicas identical, sparse and AdaRevision alike
    assert a["r1_cols"].tolist() == [7, 8] and a["r1_vals"].tolist() == [4.0, 4.0]      # column 500 went +1 -1 +1 -1 on both: gone
    assert a["r2_dense"][3] == 16.0 and a["r2_dense"][999] == 4.0
    assert np.all(a["ada"][:2] != 0) and a["ada"][0] < 0 < a["ada"][1]        # gradient +1 moves down, negative gradients up
    assert np.allclose(a["half"], 4 * (0.1 + 0.2), atol=2e-3)                  # fp16 on the wire, fp32 at rest


def test_thread_cache_and_threads():
    """ThreadInc buffers per calling thread until FlushThreadCache / the clock; RegisterThread hands out ids."""
    import threading
    from poseidon_b200.ps.table import PSTableGroup
    g = PSTableGroup.init(None, staleness=0)
    t = g.create_table(0, num_rows=1, row_capacity=4)
    g.create_table_done()
    ids = []

    def work(k):
        ids.append(g.register_thread())
        t.thread_inc(0, k, 1.0 + k)
        assert t.thread_get(0)[k].item() == 1.0 + k and t.get(0)[k].item() == 0.0      # visible to this thread only
        t.flush_thread_cache()
        g.deregister_thread()
    th = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert sorted(ids) == [0, 1, 2] and t.get(0).tolist() == [1.0, 2.0, 3.0, 0.0]
    g.shut_down()
