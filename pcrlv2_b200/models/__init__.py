"""Mirror of the reference ``models`` package (reference models/__init__.py): the 3-D and the 2-D model."""
from .pcrlv2_model_3d import PCRLv23d  # noqa: F401
from .pcrlv2_model import PCRLv2  # noqa: F401
