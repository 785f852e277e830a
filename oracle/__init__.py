"""CPU oracle for the SMOW-Net alignment/fusion hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package ``smow_net_b200`` never does.

* ``oracle.c_oracle``   numpy front-end of the plain-C restatement (oracle/smow_oracle.c)
* ``oracle.torch_ref``  PyTorch restatement that issues the very ATen calls the reference issues
                        (F.grid_sample / F.interpolate / torch.cat), for same-device comparisons
* ``oracle.make_golden`` script that ran the *real* reference here and wrote tests/golden/
"""
