"""B200-native SurroGlas per-timestep hot path behind the reference's Python API.

    from fem_glass_tempering_b200 import ThermoViscoProblem, ThermalModel, ViscoelasticModel, create_mesh

The package directory uses an underscore (a hyphen is not importable in Python); it is the
`fem-glass-tempering_b200` package of the task statement.
"""
from .models import ThermalModel, ViscoelasticModel, prony_tables  # noqa: F401
from .problem import ThermoViscoProblem  # noqa: F401
from .geometry import create_mesh  # noqa: F401
