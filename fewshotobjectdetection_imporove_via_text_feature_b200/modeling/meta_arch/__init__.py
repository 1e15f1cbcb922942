from .gdl import AffineLayer, GradientDecoupleLayer, decouple_layer, decoupled_affine
