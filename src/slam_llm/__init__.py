"""slam_llm — the reference's recipe surface (model_factory / setup_* / slam_model / finetune.main / train) over
the B200 kernels in slam_llm_b200.  Recipes from /root/reference/examples/asr_* import this package unchanged
(PYTHONPATH=<repo>/src:<repo>)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)  # so that `slam_llm_b200` (repo root) is importable when only <repo>/src is on PYTHONPATH

from ._compat import install as _install_compat  # noqa: E402

_install_compat()
