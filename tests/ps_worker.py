import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poseidon_b200 import init_rank_context
from poseidon_b200.ps import PSTableGroup

rc = init_rank_context("cpu")
staleness = int(sys.argv[2])
g = PSTableGroup.init(rc, staleness=staleness)
t = g.create_table(0, num_rows=4, row_capacity=8)
g.create_table_done()
for it in range(5):
    t.inc(1, 2, float(rc.rank + 1))
    t.batch_inc(2, {0: 1.0, 3: 2.0 * (rc.rank + 1)})
    t.dense_batch_inc(3, torch.ones(8) * (it + 1))
    # read-my-writes: own increments are visible immediately
    assert t.get(1)[2].item() >= (it + 1) * (rc.rank + 1)
    g.clock()
    if staleness == 0:
        # BSP: after clock c every worker's updates of clocks < c are visible
        w = rc.world_size
        assert abs(t.get(1)[2].item() - (it + 1) * w * (w + 1) / 2) < 1e-5
g.global_barrier()
np.save(f"{sys.argv[1]}.{rc.rank}.npy", t.data.numpy())
g.shut_down()
rc.shutdown()
