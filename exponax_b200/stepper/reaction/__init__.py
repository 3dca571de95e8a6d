from ._polynomial_reactions import AllenCahn, FisherKPP, SwiftHohenberg

__all__ = ["AllenCahn", "FisherKPP", "SwiftHohenberg"]
