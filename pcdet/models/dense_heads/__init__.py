from cagroup3d_b200.head import CAGroup3DHead

__all__ = {"CAGroup3DHead": CAGroup3DHead}
