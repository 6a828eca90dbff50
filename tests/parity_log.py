"""Collects the measured parity margins of the GPU tests (see tests/conftest.py::pytest_sessionfinish)."""
import json
import os

_ROWS = []


def _current_test():
    return os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]


def record(quantity, measured, tol, **extra):
    """quantity: what was compared (e.g. 'Z', 'loss', 'grad:conv_final1.weight'); measured: the error; tol: the bound."""
    row = {"test": _current_test(), "quantity": quantity, "measured": float(measured), "tol": float(tol)}
    row.update(extra)
    _ROWS.append(row)


def dump(directory):
    if not _ROWS or not os.path.isdir(directory):
        return
    path = os.path.join(directory, "parity.json")
    rows = []
    if os.path.isfile(path):           # several pytest invocations in one gpurun call append to the same table
        try:
            rows = json.load(open(path))
        except Exception:
            rows = []
    seen = {(r["test"], r["quantity"]) for r in _ROWS}
    rows = [r for r in rows if (r["test"], r["quantity"]) not in seen] + _ROWS
    with open(path, "w") as f:
        json.dump(rows, f, indent=0)
