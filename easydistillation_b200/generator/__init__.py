from .elemental import ElementalGenerator
from .displacement_elemental import DisplacementElementalGenerator
from .laplacian import Laplacian
