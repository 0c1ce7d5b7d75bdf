"""NumPy restatement of the loader's per-graph standardisation (TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``).

``processing/data.py:467-506``: ``StandardScaler().fit(X); X = scaler.transform(X)`` on the node features (columns 1..
when ``regularization.cell_type`` names column 0) and on the edge features.  sklearn 1.0.1 semantics restated: mean and
population variance per column accumulated in float64, ``scale_ = sqrt(var_)`` with (near-)constant columns set to 1
(``_handle_zeros_in_scale``), output in the input's float32.  ``tests/test_loader_cpu.py`` pins it to the installed
sklearn.
"""
import numpy as np


def standardize(x: np.ndarray, skip_first: bool = False) -> np.ndarray:
    out = np.array(x, dtype=np.float32, copy=True)
    sub = out[:, 1:] if skip_first else out
    x64 = sub.astype(np.float64)
    n = x64.shape[0]
    mean = x64.sum(axis=0) / n
    d = x64 - mean
    var = ((d * d).sum(axis=0) - d.sum(axis=0) ** 2 / n) / n
    scale = np.sqrt(np.maximum(var, 0.0))
    const = var <= 10 * np.finfo(np.float64).eps * n * mean * mean
    scale[const | (scale == 0)] = 1.0
    sub[:] = ((x64 - mean) / scale).astype(np.float32)
    return out
