"""B200-native hydroelastic contact-surface engine behind the mujoco_contact_surfaces plugin surface.

Only what the hot path needs lives here: csrc/ (hand-written CUDA for sm_100a + the C ABI of
include/hcs.h), engine.py (ctypes binding), scenes.py (synthetic scenes of the BASELINE shapes),
sharding.py (env-index sharding across GPUs), plugin/ (C++ host adapter mirroring the reference classes).
"""
from .engine import (GEOM_BOX, GEOM_CYLINDER, GEOM_ELLIPSOID, GEOM_MESH, GEOM_PLANE, GEOM_SPHERE,  # noqa: F401
                     REP_POLYGON, REP_TRIANGLE, WINDOW_GAUSS, WINDOW_NONE, WINDOW_SQUARE, WINDOW_TUKEY,
                     HcsError, HydroelasticEngine, MultiDeviceEngine, load_library, version)
