"""Mirror of the reference ``models`` package for the 3-D path (reference models/__init__.py:2).
The 2-D ``PCRLv2`` model is outside the scope of this build ("next")."""
from .pcrlv2_model_3d import PCRLv23d  # noqa: F401


def __getattr__(name):
    if name == "PCRLv2":
        raise NotImplementedError("the 2-D PCRLv2 model is not part of the B200 3-D hot path build")
    raise AttributeError(name)
