"""u2mkd_b200 — B200-native (sm_100a) LiDAR point-voxel backbone hot path of U2MKD."""
__version__ = "0.1.0"
