from oracle.upstream import Registry, build_fcos_resnet_fpn_backbone
BACKBONE_REGISTRY = Registry("BACKBONE")
BACKBONE_REGISTRY.register(build_fcos_resnet_fpn_backbone)


def build_backbone(cfg, input_shape=None):
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)
