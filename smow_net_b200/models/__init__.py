"""Drop-in SMOW-Net modules: ``from smow_net_b200.models.SMOW_Net import SMOW_Net`` (or put the
package directory on sys.path to keep the reference's ``from models.SMOW_Net import SMOW_Net``)."""
from .SMOW_Net import SMOW_Net          # noqa: F401
from .SMOW_Net_LW import SMOW_Net_LW    # noqa: F401
