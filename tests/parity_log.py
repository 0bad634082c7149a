"""Every number a GPU parity test measures is also written down: ``record(test, **numbers)`` collects
them and the session writes ``gpurun_out/r02_parity.json`` (the directory a GPU-box run brings back);
the copy under ``profiles/`` is the tracked one."""
import atexit
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "gpurun_out", "r02_parity.json")
_LOG = {}


def record(test: str, **numbers) -> None:
    entry = _LOG.setdefault(test, {})
    for k, v in numbers.items():
        entry[k] = float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v


def _flush() -> None:
    if not _LOG:
        return
    os.makedirs(os.path.dirname(PATH), exist_ok=True)
    merged = {}
    try:
        with open(PATH) as f:
            merged = json.load(f)
    except Exception:
        pass
    merged.update(_LOG)
    with open(PATH, "w") as f:
        json.dump(merged, f, indent=1, sort_keys=True)


atexit.register(_flush)
