"""B200-native hot path of tteepe/CenterNet-pytorch-lightning (see DESIGN.md / INTEGRATION.md).

Sub-packages mirror the reference's import sites name for name: `models` (create_model, heads.CenterHead),
`DCN.dcn_v2` (DCN), `decode.ctdet` / `decode.multi_pose`, `utils.losses`, `utils.decode`.
"""
import sys

from . import DCN, decode, models, utils
from .DCN import dcn_v2  # noqa: F401
from .decode import ctdet, multi_pose  # noqa: F401
from .models import heads  # noqa: F401
from .models.exec_modes import precision, set_precision  # noqa: F401
from .utils import decode as _utils_decode, losses as _utils_losses  # noqa: F401

SWAP = {
    "DCN": DCN, "DCN.dcn_v2": DCN.dcn_v2,
    "CenterNet.models": models, "CenterNet.models.heads": models.heads,
    "CenterNet.decode.ctdet": decode.ctdet, "CenterNet.decode.multi_pose": decode.multi_pose,
    "CenterNet.utils.losses": utils.losses, "CenterNet.utils.decode": utils.decode,
}


def install_swap():
    """The reference-side import swap of INTEGRATION.md section 1 as one call: after it, the reference's
    `centernet.py` / `centernet_detection.py` / `centernet_multi_pose.py` import this package's modules wherever they
    import `CenterNet.models`, `CenterNet.utils.losses`, `DCN.dcn_v2`, ... .  Returns the previous entries."""
    saved = {k: sys.modules.get(k) for k in SWAP}
    sys.modules.update(SWAP)
    return saved
