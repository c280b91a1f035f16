"""Two workers on a sparse (sorted-vector-map) table and an AdaRevision table; every replica must end identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poseidon_b200 import init_rank_context
from poseidon_b200.ps import AdaRevisionServerTableLogic, PSTableGroup, rows

rc = init_rank_context("cpu")
g = PSTableGroup.init(rc, staleness=0)
counts = g.create_table(0, num_rows=3, row_capacity=1000, row_type=rows.SORTED_VECTOR_MAP_ROW)
ada = g.create_table(1, num_rows=2, row_capacity=4, row_type=rows.DENSE_FLOAT_ROW,
                     table_logic=AdaRevisionServerTableLogic(init_step_size=0.1))
half = g.create_table(2, num_rows=1, row_capacity=8, row_type=rows.DENSE_FLOAT16_ROW)
g.create_table_done()
for it in range(4):
    # LDA-style counts: +1 on a topic column, -1 on another (entries that return to zero disappear)
    counts.inc(1, 7 + rc.rank, 1.0)
    counts.inc(1, 500, 1.0 if it % 2 == 0 else -1.0)
    counts.batch_inc(2, {3: 2.0, 999: float(rc.rank)})
    assert counts.get(1)[7 + rc.rank].item() >= it + 1          # read-my-writes on logic-free tables
    ada.batch_inc(0, {0: 1.0, 1: -0.5 * (rc.rank + 1)})          # gradients
    half.dense_batch_inc(0, torch.full((8,), 0.1 * (rc.rank + 1)))
    g.clock()
g.global_barrier()
r1, r2 = counts.get(1), counts.get(2)
out = {"r1_cols": np.array(r1.cols.tolist()), "r1_vals": np.array(r1.vals.tolist()), "r2_dense": r2.to_dense().numpy(),
       "ada": ada.get(0).to_dense().numpy(), "half": half.get(0).to_dense().numpy(), "n_old": np.array(len(ada.logic.old_accum))}
np.savez(f"{sys.argv[1]}.{rc.rank}.npz", **out)
g.shut_down()
rc.shutdown()
