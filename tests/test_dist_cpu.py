"""Multi-process data-parallel semantics on CPU (gloo, world_size 2): summed-update all-reduce, library SFB,
SSP bounded staleness — checked against single-process training on the concatenated batch."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(world, out, extra, device="cpu", timeout=300, local_world=None):
    """``local_world``: ranks per simulated node (torchrun's LOCAL_WORLD_SIZE / LOCAL_RANK layout, node-major)."""
    port = _free_port()
    procs = []
    lw = local_world or world
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r % lw), LOCAL_WORLD_SIZE=str(lw), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        cmd = [sys.executable, os.path.join(HERE, "dist_worker.py"), "--out", out] + extra
        if device:
            cmd += ["--device", device]
        procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return [dict(np.load(f"{out}.{r}.npz")) for r in range(world)]


def _weights(d):
    return {k: v for k, v in d.items() if "." in k and not k.startswith("wire_")}


def _assert_close(a, b, tol=2e-5):
    for k in _weights(a):
        assert np.allclose(a[k], b[k], atol=tol, rtol=1e-4), (k, np.abs(a[k] - b[k]).max())


@pytest.fixture(scope="module")
def single(tmp_path_factory):
    """1 process, batch 2M, lr x2  ==  2 workers with summed updates (reference semantics)."""
    out = str(tmp_path_factory.mktemp("single") / "w")
    return launch(1, out, ["--batch", "16", "--base_lr", "0.02"])[0]


def test_dense_allreduce_sum_semantics(tmp_path, single):
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "gloo"])
    _assert_close(res[0], res[1], 1e-7)           # replicas identical
    _assert_close(res[0], single)
    assert int(res[0]["wire_dense_allreduce_bytes"]) > 0


def test_grad_reduce_mean_equals_large_batch(tmp_path):
    ref = launch(1, str(tmp_path / "s"), ["--batch", "16", "--base_lr", "0.01"])[0]
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "gloo", "--grad_reduce", "mean"])
    _assert_close(res[0], ref)


def test_sufficient_factor_broadcast_equals_dense(tmp_path, single):
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "gloo", "--svb", "1", "--sfb_mode", "all"])
    _assert_close(res[0], res[1], 1e-6)
    _assert_close(res[0], single, 5e-5)


def test_ssp_staleness_zero_is_bsp_with_summed_updates(tmp_path, single):
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "ssp", "--staleness", "0"])
    _assert_close(res[0], res[1], 1e-6)
    _assert_close(res[0], single, 5e-5)


@pytest.mark.parametrize("comm", ["ssp", "ssp_aggr"])
def test_ssp_with_sufficient_factors_at_staleness_zero_is_bsp(tmp_path, single, comm):
    """--svb=true with a bounded-staleness backend (the caffe_main usage example): SFB weights carry the GLOBAL gradient,
    so they are stepped locally and never exchanged again (round-1 advisor finding: they used to get world x the update)."""
    extra = ["--aggr_fraction", "1.0"] if comm == "ssp_aggr" else []
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", comm, "--staleness", "0", "--svb", "1", "--sfb_mode", "all"]
                 + extra)
    _assert_close(res[0], res[1], 1e-6)
    _assert_close(res[0], single, 5e-5)


def test_ssp_bounded_staleness_converges_to_same_table(tmp_path):
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "ssp", "--staleness", "2", "--steps", "6",
                                            "--delay_rank", "1"])
    # every update is applied exactly once everywhere: after the final drain all replicas hold the same table
    _assert_close(res[0], res[1], 1e-5)
    assert int(res[0]["max_lag"]) <= 2 and int(res[1]["max_lag"]) <= 2
    assert np.isfinite(res[0]["loss"])


def test_injected_straggler_does_not_change_bsp_result(tmp_path, single, monkeypatch):
    """POSEIDON_FAULT delay on one rank (fault-injection hook): BSP training waits for the straggler and
    produces exactly the same weights."""
    monkeypatch.setenv("POSEIDON_FAULT", "delay:rank=1,step=1,ms=300")
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "gloo"])
    _assert_close(res[0], res[1], 1e-7)
    _assert_close(res[0], single)


def test_ssp_aggr_full_fraction_is_bsp(tmp_path, single):
    """SSPAggr sending the whole residual every clock = BSP with summed per-worker updates."""
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "ssp_aggr", "--aggr_fraction", "1.0"])
    _assert_close(res[0], res[1], 1e-6)
    _assert_close(res[0], single, 5e-5)


def test_ssp_aggr_budgeted_updates_are_exactly_once_and_cheaper(tmp_path):
    """10 % of the pending update per clock (largest relative magnitude first), flush bound = staleness 2: after the final
    drain every update has been applied exactly once everywhere, and far fewer bytes crossed the wire."""
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "ssp_aggr", "--aggr_fraction", "0.1", "--staleness", "2",
                                            "--steps", "7"])
    _assert_close(res[0], res[1], 1e-5)
    assert np.isfinite(res[0]["loss"])
    sent, dense = int(res[0]["wire_ssp_aggr_bytes"]), int(res[0]["wire_dense_equiv_bytes"])
    assert 0 < sent < 0.75 * dense
    # the same schedule with nothing held back gives the BSP result; the budgeted run must stay close to it after the drain
    ref = launch(2, str(tmp_path / "r"), ["--batch", "8", "--comm", "ssp_aggr", "--aggr_fraction", "1.0", "--steps", "7"])
    diffs = [np.abs(res[0][k] - ref[0][k]).max() for k in _weights(ref[0])]
    scale = max(np.abs(v).max() for v in _weights(ref[0]).values())
    assert max(diffs) < 0.2 * scale


# ------------------------------------------------------------------------------------------------------------------
# The sm100 engine (bf16 operands, fused epilogues, shadows) under the LIBRARY communication backends — the
# configuration multi-node jobs run in.  The kernels are replaced by their CPU emulation (ops/emulate.py), the
# communication is real (gloo, 2 processes).
@pytest.fixture(scope="module")
def single_sm100(tmp_path_factory):
    os.environ["POSEIDON_EMULATE"] = "1"
    try:
        out = str(tmp_path_factory.mktemp("single_sm100") / "w")
        return launch(1, out, ["--batch", "16", "--base_lr", "0.02", "--engine", "sm100", "--comm", "local"])[0]
    finally:
        os.environ.pop("POSEIDON_EMULATE", None)


def _rel(a, b):
    return max(np.abs(a[k] - b[k]).max() / (np.abs(a[k]).max() + 1e-6) for k in _weights(a))


@pytest.mark.parametrize("extra,sfb", [(["--comm", "gloo"], False),
                                       (["--comm", "gloo", "--svb", "1", "--sfb_mode", "all"], True)])
def test_sm100_engine_with_library_backend(tmp_path, monkeypatch, single_sm100, extra, sfb):
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--engine", "sm100"] + extra)
    assert _rel(res[0], res[1]) < 1e-6                    # replicas identical
    assert _rel(res[0], single_sm100) < 1e-4              # == one process on the concatenated batch with lr x2
    if sfb:
        assert int(res[0]["wire_sfb_bytes"]) > 0 and int(res[0]["wire_sfb_dense_equiv_bytes"]) > int(res[0]["wire_sfb_bytes"])
    else:
        assert int(res[0]["wire_dense_allreduce_bytes"]) > 0


@pytest.mark.parametrize("comm", ["ssp", "ssp_aggr"])
def test_sm100_engine_bounded_staleness(tmp_path, monkeypatch, single_sm100, comm):
    """SSP / SSPAggr exchange dense wire buffers whatever the parameter's memory format is (channels-last conv
    weights on this engine) and the bf16 operands are re-derived after remote updates are applied."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--engine", "sm100", "--comm", comm, "--staleness", "1"])
    assert _rel(res[0], res[1]) < 1e-5                    # drained at the end: replicas agree
    assert _rel(res[0], single_sm100) < 0.25              # and stay near the synchronous trajectory
    if "max_lag" in res[0]:
        assert int(res[0]["max_lag"]) <= 1


# ------------------------------------------------------------------------------------------------------------------
# The fused NVLink backend itself, ranks as processes: the symmetric arena is host shared memory, the peer ops are the
# emulations of ops/emulate.py (same pointers-into-the-arena calling convention, same epoch / flag protocol as the
# kernels).  What runs for real is parallel/fused.py: arena carving, re-homing of masters / histories / bf16 shadows,
# gradient sinks, DWBP bucket launches, one- vs two-shot selection, SFB slots with consumed flags, epoch bookkeeping.
@pytest.mark.parametrize("extra,sfb", [(["--comm", "fused"], False),
                                       (["--comm", "fused", "--svb", "1", "--sfb_mode", "all"], True)])
def test_fused_backend_two_ranks_on_emulated_peer_memory(tmp_path, monkeypatch, single_sm100, extra, sfb):
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--engine", "sm100"] + extra)
    assert _rel(res[0], res[1]) == 0.0                     # every rank applies the same bits
    assert _rel(res[0], single_sm100) < 1e-6               # == one process on the concatenated batch with lr x2
    if sfb:
        assert int(res[0]["wire_sfb_bytes"]) > 0
        assert int(res[0]["wire_sfb_dense_equiv_bytes"]) > 8 * int(res[0]["wire_sfb_bytes"])
    else:
        assert int(res[0]["wire_sfb_bytes"]) == 0 and int(res[0]["wire_dense_allreduce_bytes"]) > 0


def test_fused_ssp_staleness_zero_is_bsp_with_summed_updates(tmp_path, monkeypatch, single_sm100):
    """The fused engine's SSP kernels (ssp_delta / ssp_fold on the arena's delta rings, no library collective) at
    staleness 0: every worker's delta of clock c is folded before clock c + 1 starts = BSP with summed per-worker updates
    = one process on the concatenated batch with lr x 2."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    monkeypatch.setenv("POSEIDON_FUSED_SSP", "1")
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--engine", "sm100", "--comm", "fused", "--staleness", "0"])
    lib = launch(2, str(tmp_path / "l"), ["--batch", "8", "--engine", "sm100", "--comm", "ssp", "--staleness", "0"])
    assert _rel(res[0], res[1]) < 1e-6
    assert _rel(res[0], lib[0]) < 1e-6                # the library SSP backend is the oracle: same update, same order of magnitude of rounding
    assert _rel(res[0], single_sm100) < 2e-2          # (w - d_p) - d_q vs w - (d_p + d_q): last-bit differences that bf16 operands amplify
    assert int(res[0]["max_lag"]) == 0 and int(res[0]["wire_ssp_delta_bytes"]) > 0


@pytest.mark.parametrize("world,staleness", [(2, 1), (3, 2)])
def test_fused_ssp_bounded_staleness_with_a_straggler(tmp_path, monkeypatch, single_sm100, world, staleness):
    """A delayed rank (POSEIDON_FAULT) under the fused SSP kernels: nobody observes a peer more than `staleness` clocks
    behind, every delta is folded exactly once everywhere (after the final drain all replicas hold the same table), and
    the trajectory stays near the synchronous one.  Same contract as the library SSP backend."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    monkeypatch.setenv("POSEIDON_FAULT", "delay:rank=1,step=2,ms=400")
    res = launch(world, str(tmp_path / "w"), ["--batch", "8", "--engine", "sm100", "--comm", "fused", "--staleness",
                                               str(staleness), "--steps", "6"])
    for r in res[1:]:
        assert _rel(res[0], r) < 1e-5
    assert all(int(r["max_lag"]) <= staleness for r in res)
    assert np.isfinite(res[0]["loss"])
    lib = launch(world, str(tmp_path / "l"), ["--batch", "8", "--engine", "sm100", "--comm", "ssp", "--staleness", str(staleness),
                                               "--steps", "6"])
    # both are async trajectories of the same job whose fold order depends on timing: same neighbourhood, not equal
    # (0.25 was observed on a loaded machine; the strict invariants are the replica equality and the lag bound above)
    assert _rel(res[0], lib[0]) < 0.4


def test_fused_backend_three_ranks_two_shot_equals_library_backend(tmp_path, monkeypatch):
    """Sharded (two-shot) reduce + step + broadcast with a remainder shard (3 ranks), SFB on: same weights as the gloo
    all-reduce backend on the same engine."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    lib = launch(3, str(tmp_path / "g"), ["--batch", "8", "--engine", "sm100", "--comm", "gloo"])
    monkeypatch.setenv("POSEIDON_ONE_SHOT_BYTES", "1024")
    two = launch(3, str(tmp_path / "t"), ["--batch", "8", "--engine", "sm100", "--comm", "fused", "--svb", "1",
                                         "--sfb_mode", "all"])
    assert _rel(two[0], two[1]) == 0.0 and _rel(two[0], two[2]) == 0.0
    assert _rel(two[0], lib[0]) < 1e-6           # (the one-shot schedule is compared with it in the 2-rank tests)


@pytest.mark.parametrize("comm", [["--comm", "fused", "--svb", "1", "--sfb_mode", "all"], ["--comm", "gloo"]])
def test_multi_rank_snapshot_and_resume_is_exact(tmp_path, monkeypatch, comm):
    """2 steps + snapshot, then a fresh job restores and runs 2 more == 4 uninterrupted steps, bit for bit.  Covers
    the collective inside snapshot (two-shot buckets keep the history sharded by rank: every rank must join the
    gather) and the re-derivation of the bf16 operands on EVERY rank after the restored weights are broadcast."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    monkeypatch.setenv("POSEIDON_ONE_SHOT_BYTES", "1024")
    base = ["--batch", "8", "--engine", "sm100"] + comm
    full = launch(2, str(tmp_path / "A"), base + ["--steps", "4"])
    launch(2, str(tmp_path / "B"), base + ["--steps", "2", "--total_steps", "4", "--snapshot_prefix", str(tmp_path / "snap")])
    assert os.path.exists(tmp_path / "snap_iter_2.solverstate") and os.path.exists(tmp_path / "snap_iter_2.caffemodel")
    res = launch(2, str(tmp_path / "C"), base + ["--steps", "2", "--total_steps", "4", "--restore",
                                                 str(tmp_path / "snap_iter_2.solverstate")])
    assert _rel(res[0], res[1]) == 0.0
    assert _rel(res[0], full[0]) < 1e-6


def test_fused_backend_with_frozen_layers_equals_library_backend(tmp_path, monkeypatch):
    """Finetuning shape: blobs_lr 0 on a convolution and an inner product (no gradient, no bucket, no arena segment)."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    fz = ["--batch", "8", "--engine", "sm100", "--freeze", "conv2,fc4"]
    lib = launch(2, str(tmp_path / "g"), fz + ["--comm", "gloo"])
    fused = launch(2, str(tmp_path / "f"), fz + ["--comm", "fused", "--svb", "1", "--sfb_mode", "all"])
    start = launch(2, str(tmp_path / "z"), fz + ["--comm", "gloo", "--steps", "1"])
    assert _rel(fused[0], fused[1]) == 0.0 and _rel(fused[0], lib[0]) < 1e-6
    for k in ("conv2.0", "conv2.1", "fc4.0", "fc4.1"):                 # frozen blobs did not move between step 1 and 3
        assert np.array_equal(fused[0][k], start[0][k]), k
    assert not np.array_equal(fused[0]["conv3.0"], start[0]["conv3.0"])


def test_fused_backend_two_nodes_hierarchical(tmp_path, monkeypatch):
    """2 simulated nodes x 2 ranks: NVLink arena per node (node-local process group), gradients all-reduced across
    nodes by the library among ranks of equal local index, then the node-local reduce + step + broadcast kernel.
    Same weights as the flat gloo all-reduce on 4 ranks; snapshot (history gathered per node) + resume is exact."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    monkeypatch.setenv("POSEIDON_ONE_SHOT_BYTES", "1024")
    base = ["--batch", "4", "--engine", "sm100"]
    lib = launch(4, str(tmp_path / "g"), base + ["--comm", "gloo", "--steps", "4"])
    full = launch(4, str(tmp_path / "f"), base + ["--comm", "fused", "--svb", "1", "--steps", "4"], local_world=2)
    assert all(_rel(full[0], full[i]) == 0.0 for i in (1, 2, 3))
    assert _rel(full[0], lib[0]) < 1e-6
    assert int(full[0]["wire_sfb_bytes"]) == 0
    # sharded buckets: only this rank's shard (1 / GPUs-per-node of the bucket) crosses the network
    assert 0 < int(full[0]["wire_inter_node_allreduce_bytes"]) < 0.6 * int(full[0]["wire_dense_allreduce_bytes"])
    launch(4, str(tmp_path / "B"), base + ["--comm", "fused", "--steps", "2", "--total_steps", "4",
                                            "--snapshot_prefix", str(tmp_path / "snap")], local_world=2)
    res = launch(4, str(tmp_path / "C"), base + ["--comm", "fused", "--steps", "2", "--total_steps", "4", "--restore",
                                                 str(tmp_path / "snap_iter_2.solverstate")], local_world=2)
    assert _rel(res[3], full[0]) < 1e-6


def test_bf16_wire_format_halves_network_bytes(tmp_path, monkeypatch, single):
    """--wire_dtype bf16 (the reference's DenseFloat16 row oplogs): gradients cross the network as bf16, masters and
    the reduction result stay fp32.  Half the bytes, weights within bf16 rounding of the fp32-wire run; on the fused
    engine only the inter-node hop is affected."""
    fp = launch(2, str(tmp_path / "a"), ["--batch", "8", "--comm", "gloo"])
    bf = launch(2, str(tmp_path / "b"), ["--batch", "8", "--comm", "gloo", "--wire_dtype", "bf16"])
    assert int(bf[0]["wire_dense_allreduce_bytes"]) * 2 == int(fp[0]["wire_dense_allreduce_bytes"])
    assert _rel(bf[0], bf[1]) < 1e-7 and 0 < _rel(bf[0], fp[0]) < 0.02
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    base = ["--batch", "4", "--engine", "sm100", "--comm", "fused"]
    f32 = launch(4, str(tmp_path / "c"), base, local_world=2)
    f16 = launch(4, str(tmp_path / "d"), base + ["--wire_dtype", "bf16"], local_world=2)
    assert int(f16[0]["wire_inter_node_allreduce_bytes"]) * 2 == int(f32[0]["wire_inter_node_allreduce_bytes"])
    assert int(f16[0]["wire_dense_allreduce_bytes"]) == int(f32[0]["wire_dense_allreduce_bytes"])       # NVLink part
    assert all(_rel(f16[0], f16[i]) == 0.0 for i in (1, 2, 3)) and 0 < _rel(f16[0], f32[0]) < 0.02


def test_ssp_aggr_snapshot_is_per_worker(tmp_path, monkeypatch):
    """Momentum is per worker in the asynchronous modes: SSPAggr writes one .solverstate.<rank>.0 per rank like SSP (the
    reference's per-thread solverstate suffix), and a resumed job continues the uninterrupted trajectory."""
    monkeypatch.setenv("POSEIDON_EMULATE", "1")
    base = ["--batch", "8", "--engine", "sm100", "--comm", "ssp_aggr", "--staleness", "1"]
    full = launch(2, str(tmp_path / "A"), base + ["--steps", "4"])
    launch(2, str(tmp_path / "B"), base + ["--steps", "2", "--total_steps", "4", "--snapshot_prefix", str(tmp_path / "snap")])
    assert os.path.exists(tmp_path / "snap_iter_2.solverstate.0.0") and os.path.exists(tmp_path / "snap_iter_2.solverstate.1.0")
    res = launch(2, str(tmp_path / "C"), base + ["--steps", "2", "--total_steps", "4", "--restore",
                                                 str(tmp_path / "snap_iter_2.solverstate")])
    assert _rel(res[0], full[0]) < 1e-5
