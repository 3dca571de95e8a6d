"""Concrete PDE steppers (mirrors exponax/stepper/__init__.py for the hot-path scope)."""
from . import generic as generic
from . import reaction as reaction
from ._burgers import Burgers
from ._korteweg_de_vries import KortewegDeVries
from ._kuramoto_sivashinsky import KuramotoSivashinsky, KuramotoSivashinskyConservative
from ._linear import Advection, AdvectionDiffusion, Diffusion, Dispersion, HyperDiffusion
from ._wave import Wave
from ._navier_stokes import (
    KolmogorovFlowVelocity,
    KolmogorovFlowVorticity,
    NavierStokesVelocity,
    NavierStokesVorticity,
)

__all__ = [
    "Advection",
    "Diffusion",
    "AdvectionDiffusion",
    "Dispersion",
    "HyperDiffusion",
    "Wave",
    "Burgers",
    "KortewegDeVries",
    "KuramotoSivashinsky",
    "KuramotoSivashinskyConservative",
    "NavierStokesVorticity",
    "KolmogorovFlowVorticity",
    "NavierStokesVelocity",
    "KolmogorovFlowVelocity",
    "reaction",
    "generic",
]
