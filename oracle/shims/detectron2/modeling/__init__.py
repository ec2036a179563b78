from oracle.upstream import Registry
META_ARCH_REGISTRY = Registry("META_ARCH")
from .backbone import build_backbone, BACKBONE_REGISTRY  # noqa
from .proposal_generator import build_proposal_generator, PROPOSAL_GENERATOR_REGISTRY  # noqa
