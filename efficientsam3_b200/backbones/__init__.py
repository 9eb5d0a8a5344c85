from .efficientvit import (  # noqa: F401
    EfficientViTBackbone,
    efficientvit_backbone_b0,
    efficientvit_backbone_b1,
    efficientvit_backbone_b2,
)
