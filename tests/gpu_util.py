"""Helpers for the GPU parity tests (importable without a GPU)."""
import json
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(REPO, "gpurun_out", "parity_report.jsonl")

# Tolerances of BASELINE.json north_star, written once:
RTOL_GRID = 1e-5      # interpolated grids, losses, gradients (fp32)
ATOL_QP = 1e-6        # DEC q / p


def record(name, got, truth, rtol, atol):
    """Assert |got - truth| <= atol + rtol |truth| elementwise; log the achieved error."""
    got = np.asarray(got, np.float64)
    truth = np.asarray(truth, np.float64)
    assert got.shape == truth.shape, f"{name}: shape {got.shape} vs {truth.shape}"
    same_nan = np.array_equal(np.isnan(got), np.isnan(truth))
    fin = np.isfinite(truth) & np.isfinite(got)
    same_inf = np.array_equal(got[~fin & ~np.isnan(truth)], truth[~fin & ~np.isnan(truth)])
    err = np.abs(got[fin] - truth[fin])
    lim = atol + rtol * np.abs(truth[fin])
    ratio = float((err / lim).max()) if err.size else 0.0
    rel = float((err / np.maximum(np.abs(truth[fin]), 1e-30)).max()) if err.size else 0.0
    entry = dict(name=name, max_abs_err=float(err.max()) if err.size else 0.0, max_rel_err=rel,
                 worst_ratio=ratio, rtol=rtol, atol=atol, n=int(err.size))
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(entry) + "\n")
    except OSError:
        pass
    assert same_nan, f"{name}: NaN pattern differs"
    assert same_inf, f"{name}: inf pattern differs"
    assert ratio <= 1.0, f"{name}: {entry}"
    return entry


def scale_atol(truth, rtol):
    """Absolute floor for values near zero: rtol x the tensor's typical magnitude."""
    t = np.asarray(truth, np.float64)
    t = t[np.isfinite(t)]
    return rtol * float(np.sqrt(np.mean(t * t))) if t.size else rtol
