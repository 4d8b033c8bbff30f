"""B200-native elemental generation: the ElementalGenerator / DisplacementElementalGenerator
path of IHEP-LQCD/EasyDistillation behind the reference's own class API, running on
hand-written sm_100a kernels (libedk_sm100a.so).  See DESIGN.md."""
from .constant import Nc, Nd, Ns
from .generator import DisplacementElementalGenerator, ElementalGenerator, Laplacian
from .insertion.derivative import derivative
from .insertion.phase import MomentumPhase
from .preset import (
    EigenvectorDevice,
    EigenvectorHostmem,
    EigenvectorNpy,
    EigenvectorTimeSlice,
    ElementalBinary,
    ElementalNpy,
    GaugeFieldBinary,
    GaugeFieldDevice,
    GaugeFieldHostmem,
    GaugeFieldIldg,
    GaugeFieldNpy,
)

__all__ = [
    "ElementalGenerator", "DisplacementElementalGenerator", "Laplacian", "MomentumPhase", "derivative",
    "GaugeFieldBinary", "GaugeFieldNpy", "GaugeFieldHostmem", "GaugeFieldIldg", "EigenvectorNpy", "EigenvectorHostmem",
    "EigenvectorTimeSlice", "GaugeFieldDevice", "EigenvectorDevice",
    "ElementalNpy", "ElementalBinary", "Nc", "Ns", "Nd",
]
