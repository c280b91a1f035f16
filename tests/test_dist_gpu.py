"""Multi-GPU (>= 2 B200) checks of the fused NVLink backend: in-kernel all-reduce + optimizer, fused SFB,
against single-GPU training on the concatenated batch; plus the NCCL baseline."""
import numpy as np
import pytest

from test_dist_cpu import _weights, launch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _close(a, b, rel):
    for k in _weights(a):
        d = np.abs(a[k] - b[k]).max()
        m = np.abs(b[k]).max() + 1e-6
        assert d <= rel * m + 2e-4, (k, d, m)


@pytest.fixture(scope="module")
def single_sm100(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("single_gpu") / "w")
    return launch(1, out, ["--batch", "16", "--base_lr", "0.02", "--engine", "sm100", "--hw", "35"], device=None)[0]


def test_fused_allreduce_sgd_two_ranks(tmp_path, single_sm100):
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--engine", "sm100", "--comm", "fused", "--svb", "0",
                                            "--hw", "35"], device=None)
    _close(res[0], res[1], 1e-6)                     # replicas bit-close
    _close(res[0], single_sm100, 0.03)               # == one GPU on the concatenated batch (bf16 noise)
    assert int(res[0]["wire_dense_allreduce_bytes"]) > 0


def test_fused_sfb_two_ranks(tmp_path, single_sm100):
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--engine", "sm100", "--comm", "fused", "--svb", "1",
                                            "--sfb_mode", "all", "--hw", "35"], device=None)
    _close(res[0], res[1], 1e-6)
    _close(res[0], single_sm100, 0.03)
    assert int(res[0]["wire_sfb_bytes"]) > 0


def test_nccl_baseline_two_ranks(tmp_path):
    ref = launch(1, str(tmp_path / "s"), ["--batch", "16", "--base_lr", "0.02", "--hw", "35"], device=None)[0]
    res = launch(2, str(tmp_path / "w"), ["--batch", "8", "--comm", "nccl", "--hw", "35"], device=None)
    _close(res[0], ref, 1e-3)


@pytest.mark.parametrize("comm", [["--comm", "fused", "--svb", "1", "--sfb_mode", "all"], ["--comm", "nccl"],
                                  ["--comm", "ssp", "--staleness", "1"]])
def test_two_gpu_snapshot_and_resume(tmp_path, monkeypatch, comm):
    """The CPU-emulated test_multi_rank_snapshot_and_resume_is_exact on real kernels / NVLink / NCCL: 2 steps + snapshot
    (sharded history gathered by every rank), restore in a fresh job, 2 more steps == 4 uninterrupted steps.  The SSP
    case also re-runs the NCCL wire-buffer path that failed once at 8 GPUs."""
    monkeypatch.setenv("POSEIDON_ONE_SHOT_BYTES", "1024")
    base = ["--batch", "8", "--engine", "sm100", "--hw", "35"] + comm
    full = launch(2, str(tmp_path / "A"), base + ["--steps", "4"], device=None)
    launch(2, str(tmp_path / "B"), base + ["--steps", "2", "--total_steps", "4", "--snapshot_prefix",
                                            str(tmp_path / "snap")], device=None)
    state = str(tmp_path / "snap_iter_2.solverstate")
    res = launch(2, str(tmp_path / "C"), base + ["--steps", "2", "--total_steps", "4", "--restore", state], device=None)
    _close(res[0], res[1], 1e-6 if comm[1] != "ssp" else 0.05)
    _close(res[0], full[0], 0.02 if comm[1] != "ssp" else 0.2)
