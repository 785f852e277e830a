"""smow_net_b200 — B200-native (sm_100a) flow-guided bi-temporal alignment/fusion path of SMOW-Net.

    from smow_net_b200.models import SMOW_Net, SMOW_Net_LW     # drop-in nn.Modules
    from smow_net_b200 import ops                               # flow_warp / tlerp_cat operators

The CUDA extension (libsmow_b200.so, C ABI in include/smow_b200.h) is mandatory; importing the
package does not need a GPU, calling an operator does.
"""
__version__ = "0.1.0"
