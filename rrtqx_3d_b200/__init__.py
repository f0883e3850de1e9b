"""rrtqx_3d_b200 -- B200 (sm_100a) implementation of RRTQX_3D's geometric hot path.

Layers:
  _abi.py        ctypes binding of the C ABI (include/rrtqx_b200.h, librrtqx_b200.so)
  device.py      handle objects: Context, DeviceTree, SphereSet, EdgeSet, results
  structures.py  host structs with the reference's fields (JList, RRTNode, SimpleEdge, ...)
  kdtree.py      KDTree, kdInsert, kdFindNearest, kdFindWithinRange, ... (kdTree_general.jl)
  collision.py   explicitEdgeCheck, explicitPointCheck, ... (DRRT_Q.jl collision section)
  sweep.py       findPointsInConflictWithObstacle, addNewObstacle, removeObstacle
  workloads.py   seeded synthetic workloads of the BASELINE configs
  sharding.py    contiguous batch sharding over ranks + result gather

Importing the package never touches CUDA; the shared library is loaded on first use and
there is no CPU fallback (no library / no B200 -> exception).
"""
__version__ = "0.1.0"
