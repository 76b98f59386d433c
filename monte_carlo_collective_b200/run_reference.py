"""Run the reference's own ``python experiments.py`` flow, unchanged, on the B200 engine.

    python -m monte_carlo_collective_b200.run_reference <reference-dir> [config.yaml]
           [--workdir DIR] [--set common.n_steps=20000 ...] [--summary out.json]

What happens (experiments.py:1204-1391, config.yaml:1-38 stay as they are):

1. the reference's ``experiments.py`` is loaded from ``<reference-dir>`` with everything above its
   ``if __name__ == "__main__":`` block executed as module ``experiments`` (definitions only);
2. ``api.install(experiments)`` rebinds the hot-path names (``run_experiment``, ``metropolis_mcmc*``, ...)
   to the engine: the drivers look them up in module globals at call time, so
   ``run_beta_start_end_pairs`` / ``run_compare_beta_end`` / ``measure_min_energy_vs_N`` and the
   ``__main__`` block itself drive the GPU without a changed line;
3. the body of the ``__main__`` block is executed, in the same namespace, with ``__name__ == "__main__"``
   and the working directory holding ``config.yaml`` (``results/*.csv`` and ``figures/`` land there).

Two accommodations, neither of which edits the reference:

* matplotlib is optional: when it cannot be imported a do-nothing ``pyplot`` stands in (the plot functions
  still write their CSV files through pandas before they draw);
* ``run_compare_beta_end`` -- the experiment the stock ``config.yaml`` selects -- passes ``annealing_type=`` and
  ``init_mode=`` to ``plot_energy_histories_side_by_side``, which does not take them (experiments.py:848,
  :1012-1022): in the stock reference the run dies with a TypeError after all chains have finished and before
  anything is returned.  Here the plot call is wrapped: per-schedule mean / std curves of both board sizes are
  written to ``results/`` first, the original function is then called as the reference calls it, and its
  TypeError is reported and swallowed so that the results reach the ``__main__`` block.
"""
from __future__ import annotations

import argparse
import ast
import functools
import json
import os
import shutil
import sys
import types

import numpy as np


class _Inert:
    """Stands in for matplotlib.pyplot: any attribute, call, index or unpacking yields another inert object."""

    def __getattr__(self, _name):
        return self

    def __call__(self, *a, **kw):
        return self

    def __getitem__(self, _k):
        return self

    def __iter__(self):
        return iter((_Inert(), _Inert()))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def ensure_pyplot():
    """Real matplotlib (Agg) when installed, else the inert stand-in.  Returns True when plots are real."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot  # noqa: F401
        return True
    except Exception:
        stub = types.ModuleType("matplotlib")
        stub.pyplot = _Inert()
        stub.use = lambda *a, **kw: None
        sys.modules["matplotlib"] = stub
        sys.modules["matplotlib.pyplot"] = stub.pyplot
        return False


def load_experiments(reference_dir):
    """(module, code of the ``__main__`` block) of ``<reference_dir>/experiments.py``."""
    path = os.path.join(reference_dir, "experiments.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    main_body, defs = None, []
    for node in tree.body:
        is_main = (isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Name)
                   and node.test.left.id == "__name__")
        if is_main:
            main_body = node.body
        else:
            defs.append(node)
    if main_body is None:
        raise RuntimeError(f"{path} has no `if __name__ == \"__main__\":` block")
    if reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)            # its own imports: mcmc, mcmc_board
    mod = types.ModuleType("experiments")
    mod.__file__ = path
    sys.modules["experiments"] = mod
    exec(compile(ast.Module(body=defs, type_ignores=[]), path, "exec"), mod.__dict__)
    return mod, compile(ast.Module(body=main_body, type_ignores=[]), path, "exec")


def _tolerant_side_by_side(original, results_dir="results"):
    """See the module docstring: results first, then the reference's own call, its TypeError survived."""
    from . import reports

    @functools.wraps(original)
    def wrapper(hist_n1, hist_n2, N1, N2, *args, **kwargs):
        for n, hists in ((N1, hist_n1), (N2, hist_n2)):
            for label, rows in hists.items():
                if len({len(r) for r in rows}) == 1:
                    arr = np.asarray(rows, dtype=np.float64)
                    reports.write_energy_csv(f"N{n}_{label}", arr.mean(axis=0), arr.std(axis=0), results_dir)
        try:
            return original(hist_n1, hist_n2, N1, N2, *args, **kwargs)
        except TypeError as exc:
            print(f"[run_reference] the reference's plot call failed as it does in the stock code ({exc}); "
                  f"results are in {results_dir}/ and are returned to the caller", file=sys.stderr)
            return None
    return wrapper


def _apply_overrides(config, overrides):
    import yaml
    for item in overrides:
        key, _, value = item.partition("=")
        node = config
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = yaml.safe_load(value)
    return config


def _summary(ns):
    """The result variables the ``__main__`` block leaves behind, JSON-friendly."""
    def conv(x):
        if isinstance(x, dict):
            return {str(k): conv(v) for k, v in x.items() if k not in ("all_histories", "all_accepted", "all_rejected")}
        if isinstance(x, (list, tuple)):
            return [conv(v) for v in x]
        if isinstance(x, np.ndarray):
            return x.tolist() if x.size <= 4096 else {"shape": list(x.shape), "mean": float(x.mean())}
        if isinstance(x, (np.integer,)):
            return int(x)
        if isinstance(x, (np.floating,)):
            return float(x)
        return x if isinstance(x, (int, float, str, bool, type(None))) else repr(type(x))
    out = {"experiment_type": ns.get("experiment_type")}
    for name in ("best_energies", "all_best_energies_dict", "steps_to_best", "result_dict"):
        if name in ns:
            out[name] = conv(ns[name])
    return out


def run(reference_dir, config_path=None, workdir=None, overrides=(), summary_path=None, install=True):
    """Programmatic form of the command line; returns the namespace the ``__main__`` block ran in."""
    import yaml
    reference_dir = os.path.abspath(reference_dir)
    config_path = os.path.abspath(config_path or os.path.join(reference_dir, "config.yaml"))
    workdir = os.path.abspath(workdir or os.getcwd())
    os.makedirs(workdir, exist_ok=True)
    target = os.path.join(workdir, "config.yaml")
    if overrides:
        with open(config_path) as f:
            cfg = _apply_overrides(yaml.safe_load(f), overrides)
        with open(target, "w") as f:
            yaml.safe_dump(cfg, f, sort_keys=False)
    elif os.path.abspath(target) != config_path:
        shutil.copyfile(config_path, target)
    ensure_pyplot()
    mod, main_code = load_experiments(reference_dir)
    if install:
        from . import api
        api.install(mod)
    if hasattr(mod, "plot_energy_histories_side_by_side"):
        mod.plot_energy_histories_side_by_side = _tolerant_side_by_side(mod.plot_energy_histories_side_by_side)
    ns = mod.__dict__
    here = os.getcwd()
    os.chdir(workdir)
    ns["__name__"] = "__main__"
    try:
        exec(main_code, ns)
    finally:
        ns["__name__"] = "experiments"
        os.chdir(here)
    if summary_path:
        with open(summary_path, "w") as f:
            json.dump(_summary(ns), f, indent=1)
    return ns


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m monte_carlo_collective_b200.run_reference", description=__doc__.split("\n\n")[0])
    ap.add_argument("reference_dir", help="checkout of galgantar/monte-carlo-collective (holds experiments.py)")
    ap.add_argument("config", nargs="?", default=None, help="config file (default: <reference_dir>/config.yaml)")
    ap.add_argument("--workdir", default=None, help="directory to run in: config.yaml is placed there, results/ and figures/ appear there")
    ap.add_argument("--set", dest="overrides", action="append", default=[], metavar="KEY=VALUE",
                    help="override a config entry, e.g. --set common.n_steps=20000 --set experiment_type=single_N")
    ap.add_argument("--summary", default=None, help="write the results left by the __main__ block to this JSON file")
    args = ap.parse_args(argv)
    run(args.reference_dir, args.config, args.workdir, args.overrides, args.summary)
    return 0


if __name__ == "__main__":
    sys.exit(main())
