"""GPU: checkpoint / resume.  A run stopped at a step and continued later -- in the same process
or from a file -- is the uninterrupted run: same histories, accept bits, bins, statistics, best states."""
import numpy as np
import pytest

import monte_carlo_collective_b200 as mcq
from monte_carlo_collective_b200 import schedules

pytestmark = pytest.mark.gpu

LIN = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}
FIELDS = ("energy_history", "accept_bits", "accept_hist", "initial_energy", "final_energy", "best_energy", "steps_to_best",
          "n_accepted", "steps_done", "final_state", "best_state")


def _same(a, b, fields=FIELDS, tag=""):
    for name in fields:
        assert (np.asarray(getattr(a, name)) == np.asarray(getattr(b, name))).all(), (tag, name)


@pytest.mark.parametrize("mode,n,algo", [("full_3d", 12, "auto"), ("board", 12, "auto"), ("board", 12, "lines"), ("full_3d", 9, "gmem"),
                                          ("board", 30, "wide"), ("board", 30, "gmem"), ("full_3d", 24, "auto")])
def test_segments_equal_one_run(engine, mode, n, algo, tmp_path):
    ns, nc = 2048, 24
    betas = schedules.beta_table(LIN, ns)
    seeds = np.arange(nc, dtype=np.uint64) * 3 + 11
    kw = dict(history="full", accept_bits=True, n_bins=50, algo=algo)
    whole = engine.run(mode, n, ns, seeds, betas, **kw)
    assert whole.step == ns and whole.record.shape == (8, nc)
    # three segments in one process, the arrays of the first carried along
    a = engine.run(mode, n, ns, seeds, betas, stop_step=640, **kw)
    assert a.step == 640 and (a.steps_done <= 640).all()
    b = engine.run(mode, n, ns, seeds, betas, resume=a, stop_step=1600, chunk_steps=256, **kw)
    c = engine.run(mode, n, ns, seeds, betas, resume=b, **kw)
    _same(whole, c, tag="in-process")
    assert (c.record == whole.record).all()
    # through a file: only states and records survive; the tail of the history must still match
    a.save_checkpoint(tmp_path / "ckpt.npz")
    ck = mcq.load_checkpoint(tmp_path / "ckpt.npz")
    d = engine.run(mode, n, ns, seeds, betas, resume=ck, **kw)
    _same(whole, d, fields=FIELDS[3:], tag="file")
    assert (d.energy_history[:, 641:] == whole.energy_history[:, 641:]).all()
    assert (d.accept_bits[:, 20:] == whole.accept_bits[:, 20:]).all()


def test_stats_and_early_stop_across_segments(engine):
    ns, nc = 1536, 32
    betas = np.stack([schedules.beta_table(LIN, ns), schedules.beta_table({"type": "constant", "beta_const": 6.0}, ns)])
    groups = (np.arange(nc) % 2).astype(np.int32)
    seeds = np.arange(nc, dtype=np.uint64) + 100
    for algo in ("table", "lines"):
        kw = dict(groups=groups, history="stats", early_stop_patience=200, algo=algo)
        whole = engine.run("board", 8, ns, seeds, betas, **kw)
        a = engine.run("board", 8, ns, seeds, betas, stop_step=512, **kw)
        b = engine.run("board", 8, ns, seeds, betas, resume=a, **kw)
        assert (whole.steps_done < ns).any()
        for name in ("stat_sum_e", "stat_sum_e2", "final_energy", "best_energy", "steps_to_best", "n_accepted", "steps_done",
                     "final_state", "best_state"):
            assert (getattr(b, name) == getattr(whole, name)).all(), (algo, name)


def test_resume_argument_errors(engine):
    ns = 256
    betas = schedules.beta_table(LIN, ns)
    seeds = np.arange(4, dtype=np.uint64)
    a = engine.run("board", 6, ns, seeds, betas, stop_step=64)
    with pytest.raises(ValueError):
        engine.run("board", 6, ns, seeds, betas, stop_step=50)             # not a multiple of 32
    with pytest.raises(ValueError):
        engine.run("board", 7, ns, seeds, betas, resume=a)                 # another problem
    with pytest.raises(ValueError):
        engine.run("board", 6, ns, seeds, betas, resume=a, stop_step=32)   # stop before start
    b = engine.run("board", 6, ns, seeds, betas, resume=a, stop_step=64)   # empty segment
    assert (b.final_state == a.final_state).all() and (b.record == a.record).all()


@pytest.mark.parametrize("mode,n", [("full_3d", 12), ("board", 12), ("board", 24)])
def test_statistics_survive_a_file_checkpoint(engine, mode, n, tmp_path):
    """A statistics run resumed from a FILE continues the sums of the earlier segments: the checkpoint carries the
    cumulative outputs, and a resumed run never continues into uninitialised arrays."""
    ns, nc = 1024, 20
    seeds = np.arange(nc, dtype=np.uint64) + 5
    kw = dict(schedules=LIN, history="stats", n_bins=16, stat_count=True)
    whole = engine.run(mode, n, ns, seeds, **kw)
    a = engine.run(mode, n, ns, seeds, stop_step=384, **kw)
    a.save_checkpoint(tmp_path / "ck.npz")
    ck = mcq.load_checkpoint(tmp_path / "ck.npz")
    assert ck.stat_sum_e is not None and ck.accept_hist is not None
    b = engine.run(mode, n, ns, seeds, resume=ck, **kw)
    for name in ("stat_sum_e", "stat_sum_e2", "stat_count", "accept_hist", "best_energy", "final_energy", "n_accepted", "final_state"):
        assert (np.asarray(getattr(b, name)) == np.asarray(getattr(whole, name))).all(), name
    # a checkpoint without the cumulative arrays (older file / want them dropped): the resumed run starts them from zero
    # for the columns it executes -- defined values, the earlier columns are simply absent
    ck.stat_sum_e = ck.stat_sum_e2 = ck.stat_count = ck.accept_hist = None
    c = engine.run(mode, n, ns, seeds, resume=ck, **kw)
    assert (np.asarray(c.stat_sum_e)[:, :385] == 0).all()
    assert (np.asarray(c.final_state) == np.asarray(whole.final_state)).all()
