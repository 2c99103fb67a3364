"""B200-native backend for the SelfPose3d / VoxelPose voxelised multi-view pose path.

See DESIGN.md for the path, its boundary and the kernels; ``selfpose3d_b200.models``
mirrors the reference's ``lib/models`` module surface.
"""
__version__ = "0.1.0"
