"""raytracercuda_b200 -- B200-native (sm_100a) acceleration-structure build + closest-hit trace behind the reference's
("Beam", d1kkop/RaytracerCuda) host interface.  The compute lives in libbeam_b200.so (hand-written CUDA, C ABI in
include/beam_b200.h); this package is the ctypes view of it plus mesh I/O and synthetic workloads.  No CPU fallback."""
from ._lib import BM_TRACE_CLOSEST, BM_TRACE_REFERENCE, BeamError, load  # noqa: F401
