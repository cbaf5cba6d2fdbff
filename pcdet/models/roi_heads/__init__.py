from cagroup3d_b200.roi_head import CAGroup3DRoIHead

__all__ = {"CAGroup3DRoIHead": CAGroup3DRoIHead}
