from cagroup3d_b200.backbone import BiResNet

__all__ = {"BiResNet": BiResNet}
