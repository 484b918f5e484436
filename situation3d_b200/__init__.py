"""situation3d_b200 -- B200-native (sm_100a) PointNet++ encoding path of SIG3D.

Drop-in for the reference's ``lib/pointnet2`` (same module and function names,
same signatures) plus the situation-conditioned token re-encoding, implemented
as hand-written CUDA behind the C ABI declared in ``include/pn2_b200.h``.

There is no CPU fallback: importing ``situation3d_b200.pointnet2._ext`` loads
``libpn2_b200.so`` and raises if it has not been built
(``python -m situation3d_b200.build``).
"""
__version__ = "0.1.0"
