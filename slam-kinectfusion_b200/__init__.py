"""slam-kinectfusion_b200 -- B200-native KinectFusion tracking-and-mapping core.

The product is the C-ABI shared library `libkfb200.so` (include/kfb200.h: hand-written
sm_100a kernels) and the C++ host facade `libkfusion_b200.so` that mirrors the reference's
`kfusion/include` classes on top of it.  This Python package is only the test/bench harness:
ctypes bindings of both libraries.  There is no CPU fallback: importing works anywhere, any
compute call requires the built extension and a Blackwell GPU and fails loudly otherwise.
"""
from .binding import (KfbError, Context, Params, Intrinsics, default_params, load_library,  # noqa: F401
                      library_path, SENSORS, exported_symbols)
from .build import build_all  # noqa: F401
from .host import (KinectFusion, HostParams, default_host_params, load_host_library, host_library_path,  # noqa: F401,E402
                   DatasetSensor, write_png_gray16, write_png_rgb8, read_intrinsics)
