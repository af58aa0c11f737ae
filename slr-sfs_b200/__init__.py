"""slr_sfs_b200 -- B200-native (sm_100a) frame-synthesis hot path of SLR-SFS.

Import name ``slr_sfs_b200`` (the directory is ``slr-sfs_b200/``; the repository
root carries a symlink so that it is importable).

Drop-in operator modules (same names and signatures as the reference's):
    slr_sfs_b200.softsplat                      <- models/softsplat.py
    slr_sfs_b200.euler_integration_manipulator  <- models/projection/euler_integration_manipulator.py
Fused joint block / clip synthesis:
    slr_sfs_b200.synthesis   (JointSplat)      slr_sfs_b200.clip (ClipRunner: a rank's frame block, group by group)
    slr_sfs_b200.level0      the unedited forward_flow call pattern on the drop-in operators
"""
from . import _lib
from . import softsplat
from . import euler_integration_manipulator
from . import synthesis
from . import sharding
from . import clip
from . import level0
from . import decoder_entry
from . import frame_sink
from .softsplat import (FunctionSoftsplat, ModuleSoftsplat, ModuleMaximumsplat,
                        ModuleMaximumWarpNormsplat)
from .euler_integration_manipulator import EulerIntegration, euler_integration
from .synthesis import JointSplat
from .clip import ClipRunner
from .frame_sink import FrameSink

__all__ = ["FunctionSoftsplat", "ModuleSoftsplat", "ModuleMaximumsplat", "ModuleMaximumWarpNormsplat",
           "EulerIntegration", "euler_integration", "JointSplat", "ClipRunner", "install_as_reference_modules"]


def install_as_reference_modules():
    """Make ``from models import softsplat`` and ``from models.projection.
    euler_integration_manipulator import ...`` (the reference's own import lines,
    models/animating_softmax_splating.py:9,26) resolve to this package, so the
    reference's model and training scripts run on it unchanged.  Call it before
    importing the reference's ``models`` package (with the reference tree on
    sys.path)."""
    import sys
    sys.modules["models.softsplat"] = softsplat
    sys.modules["models.projection.euler_integration_manipulator"] = euler_integration_manipulator
    try:
        import models  # the reference package, if importable
        models.softsplat = softsplat
        import models.projection as _proj
        _proj.euler_integration_manipulator = euler_integration_manipulator
    except Exception:
        pass
