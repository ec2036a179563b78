from .build import PROPOSAL_GENERATOR_REGISTRY, build_proposal_generator  # noqa
