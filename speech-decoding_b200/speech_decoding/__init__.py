"""Drop-in for the reference package of the same name.

Only the hot-path modules live here (models.py, utils/loss.py, utils/layout.py).  Everything else the
reference's train.py imports (speech_decoding.dataclass.*, utils.get_dataloaders, ...) is resolved from
any other `speech_decoding/` directory that follows on sys.path -- i.e. put this directory FIRST on
PYTHONPATH and the reference checkout after it; modules defined here shadow the reference's."""
import os as _os
import sys as _sys


def _extend(path, parts):
    mine = [_os.path.realpath(p) for p in path]
    for entry in list(_sys.path):
        cand = _os.path.join(entry or ".", *parts)
        if _os.path.isdir(cand) and _os.path.realpath(cand) not in mine:
            path.append(cand)
            mine.append(_os.path.realpath(cand))
    return path


__path__ = _extend(list(__path__), ["speech_decoding"])
