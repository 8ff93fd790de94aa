"""Importable alias of the `optical-flow-1_b200/` package directory (a hyphen is not a valid
Python identifier).  `import optical_flow_1_b200` resolves sub-modules from that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "optical-flow-1_b200"))

from .tvl1 import *  # noqa: F401,F403,E402
from .hs import *  # noqa: F401,F403,E402
from .occ import *  # noqa: F401,F403,E402
from . import synth, shard  # noqa: F401,E402
