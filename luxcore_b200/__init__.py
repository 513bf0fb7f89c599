"""luxcore_b200 -- B200 (sm_100a) drop-in for LuxRays' batched closest-hit intersection path.

Layout:
  csrc/      CUDA kernels + the C ABI (include/luxrays_b200.h)  -> lib/libluxrays_b200.so
  host/      C++ host layer mirroring the luxrays:: plugin surface -> lib/libluxrays_b200_host.so
  capi.py    ctypes binding of the C ABI
  scenes.py  benchmark geometry (fixtures of the reference's scenes, synthetic generators)
  rays.py    ray-batch producers
"""
__version__ = "0.1"
