"""Nonlinear functions (mirrors exponax/nonlin_fun/__init__.py)."""
from ._base import BaseNonlinearFun
from ._convection import ConvectionNonlinearFun
from ._general_nonlinear import GeneralNonlinearFun
from ._gradient_norm import GradientNormNonlinearFun
from ._leray import Leray
from ._polynomial import PolynomialNonlinearFun
from ._projected_convection import ProjectedConvection3d, ProjectedConvection3dKolmogorov
from ._vorticity_convection import VorticityConvection2d, VorticityConvection2dKolmogorov
from ._zero import ZeroNonlinearFun

__all__ = [
    "BaseNonlinearFun",
    "ConvectionNonlinearFun",
    "GeneralNonlinearFun",
    "GradientNormNonlinearFun",
    "Leray",
    "PolynomialNonlinearFun",
    "ProjectedConvection3d",
    "ProjectedConvection3dKolmogorov",
    "VorticityConvection2d",
    "VorticityConvection2dKolmogorov",
    "ZeroNonlinearFun",
]
