"""gaussianprocesspathmodelling_b200 -- B200-native (sm_100a) GP path-modelling hot path.

``GPmap`` is the drop-in module (the reference's data model + the GP entry points); ``workloads``
holds the seeded synthetic configurations of BASELINE.json; ``dist`` shards paths / grid rows /
hyper-parameter points across the GPUs of one node.  The math lives in ``libgpmap_b200.so``
(``csrc/``, C-ABI in ``include/gpmap_b200.h``); importing this package does not need a GPU, calling
the GP entry points does.
"""
from . import workloads  # noqa: F401
from . import GPmap  # noqa: F401
from .GPmap import (GPModel, check_if_valid_trajectory, export_raster, fit_gp, fit_gp_batched, lml_sweep,  # noqa: F401
                    make_theta, optimize_gp, readcsvfile, trajectories, trajectory)

__version__ = "0.1.0"
