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
