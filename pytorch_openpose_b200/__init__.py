"""B200-native OpenPose inference path: drop-in `Body` / `Hand` (reference: src/body.py, src/hand.py) over
hand-written sm_100a kernels in libopenpose_b200.so.  See DESIGN.md / INTEGRATION.md."""
from .body import Body            # noqa: F401
from .hand import Hand            # noqa: F401
from . import util                # noqa: F401


def install_as_src():
    """Register this package under the reference's module names so that `from src.body import Body`,
    `from src.hand import Hand`, `from src import util, model` (srcmx/MotionEstimation.py:12-15,
    srcmx/Batch_model.py:30-33) resolve to the B200 implementation."""
    import sys
    import types
    from . import body, hand, util as _util, model
    pkg = types.ModuleType("src")
    pkg.__path__ = []
    pkg.body, pkg.hand, pkg.util, pkg.model = body, hand, _util, model
    sys.modules.update({"src": pkg, "src.body": body, "src.hand": hand, "src.util": _util, "src.model": model})
    return pkg
