"""ETDRK integrators (mirrors exponax/etdrk/__init__.py)."""
from ._base_etdrk import BaseETDRK
from ._etdrk_0 import ETDRK0
from ._etdrk_1 import ETDRK1
from ._etdrk_2 import ETDRK2
from ._etdrk_3 import ETDRK3
from ._etdrk_4 import ETDRK4
from ._utils import roots_of_unity

__all__ = ["BaseETDRK", "ETDRK0", "ETDRK1", "ETDRK2", "ETDRK3", "ETDRK4", "roots_of_unity"]
