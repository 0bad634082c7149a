"""Run one of the reference's scripts on this package:

    python -m unet.dropin /path/to/unet-segment-pytorch/scripts/train.py --config configs/lung_tumor.yaml
    python -m unet.dropin /path/to/unet-segment-pytorch/scripts/predict.py --weights best.pt --source img.png

The reference's scripts put their own checkout first on ``sys.path`` (scripts/train.py:25-26), so a
plain ``python scripts/train.py`` always imports the reference's ``unet``.  This launcher imports
THIS ``unet`` first (it then owns ``sys.modules['unet']``), attaches the script's checkout as the
fall-through target (``unet.overlay``) and executes the script unchanged as ``__main__``: models,
losses, metrics and ``ModelEMA`` are the B200 ones; ``unet.data.dataset``, ``unet.utils.callbacks``,
``unet.utils.plots``, ``set_seed`` / ``get_device`` / ``load_config`` / ``increment_path`` are the
reference's own files.
"""
from __future__ import annotations

import os
import runpy
import sys


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        raise SystemExit(0 if argv else 2)
    script = os.path.abspath(argv[0])
    import unet   # this package, before the script can insert its own checkout in front of it
    from unet import overlay

    checkout = os.environ.get("UNET_REFERENCE_ROOT") or os.path.dirname(os.path.dirname(script))
    overlay.attach(checkout)
    import unet.data  # noqa: F401  (packages whose search path the overlay extends)
    import unet.models  # noqa: F401
    import unet.utils  # noqa: F401
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
