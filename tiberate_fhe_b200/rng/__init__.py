from .csprng import Csprng, build_cdt_tree  # noqa: F401
