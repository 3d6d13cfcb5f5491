"""catch_b200: sm_100a implementation of CATCH's probe-coverage + set-cover hot path.

Host side mirrors the reference's plugin interface (catch.filter.BaseFilter subclasses) and
calls libcatchb200.so through ctypes.  There is no CPU fallback: importing the filters works
anywhere, but running them needs the CUDA library and a B200.
"""
__version__ = '0.1.0'
