from ._cahn_hilliard import CahnHilliard
from ._gray_scott import GrayScott
from ._polynomial_reactions import AllenCahn, FisherKPP, SwiftHohenberg

__all__ = ["AllenCahn", "CahnHilliard", "FisherKPP", "GrayScott", "SwiftHohenberg"]
