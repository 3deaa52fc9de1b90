"""gravhopper_b200 -- B200-native (sm_100a) gravity engine behind GravHopper's Python API.

    from gravhopper_b200 import Simulation, IC      # as `from gravhopper import Simulation, IC`
    from gravhopper_b200 import grav                # as `gravhopper.grav` (= jbgrav)

Mirrors /root/reference/gravhopper/__init__.py:3-4.  The force backend, the jbgrav dispatch layer
and the Simulation.run leapfrog loop are hand-written CUDA behind a C ABI
(include/gravhopper_b200.h, gravhopper_b200/csrc/); there is no CPU fallback.
"""
from .gravhopper import (Simulation, IC, GravHopperException, UninitializedSimulationException,
                         ICException, UnknownAlgorithmException, ExternalPackageException,
                         force_centers)
from . import jbgrav as grav
from . import jbgrav, _jbgrav, units, potentials, ic_gpu

__version__ = "0.1.0"
__all__ = ["Simulation", "IC", "grav", "jbgrav", "_jbgrav", "units", "potentials", "GravHopperException",
           "UninitializedSimulationException", "ICException", "UnknownAlgorithmException",
           "ExternalPackageException", "force_centers"]
