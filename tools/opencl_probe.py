"""Is there an OpenCL platform on the GPU box?  (if the NVIDIA driver's OpenCL ICD is
mounted there, the reference's .cl kernel text could be JIT-run on the B200 to pin the oracle.)
Prints what it finds; never fails."""
import ctypes
import ctypes.util
import glob
import os

print("vendors dir:", os.listdir("/etc/OpenCL/vendors") if os.path.isdir("/etc/OpenCL/vendors") else None)
for pat in ("/usr/lib/x86_64-linux-gnu/libnvidia-opencl*", "/usr/lib64/libnvidia-opencl*", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL*",
            "/usr/lib/x86_64-linux-gnu/libOpenCL*"):
    print(pat, "->", glob.glob(pat))

cl = None
for name in ("libOpenCL.so.1", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1", ctypes.util.find_library("OpenCL") or "libOpenCL.so"):
    try:
        cl = ctypes.CDLL(name)
        print("loaded", name)
        break
    except OSError as e:
        print("cannot load", name, e)

if cl is not None:
    n = ctypes.c_uint(0)
    rc = cl.clGetPlatformIDs(0, None, ctypes.byref(n))
    print("clGetPlatformIDs rc", rc, "platforms", n.value)
    if rc == 0 and n.value > 0:
        plats = (ctypes.c_void_p * n.value)()
        cl.clGetPlatformIDs(n.value, plats, None)
        for p in plats:
            buf = ctypes.create_string_buffer(256)
            for what, code in (("name", 0x0902), ("vendor", 0x0903), ("version", 0x0901)):
                cl.clGetPlatformInfo(ctypes.c_void_p(p), code, 256, buf, None)
                print("  platform", what, buf.value.decode())
            nd = ctypes.c_uint(0)
            rc = cl.clGetDeviceIDs(ctypes.c_void_p(p), ctypes.c_uint64(0xFFFFFFFF), 0, None, ctypes.byref(nd))
            print("  clGetDeviceIDs rc", rc, "devices", nd.value)

# try loading the NVIDIA ICD directly (no vendors file needed for this)
for lib in glob.glob("/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so*") + glob.glob("/usr/lib64/libnvidia-opencl.so*"):
    try:
        h = ctypes.CDLL(lib)
        print("direct load ok:", lib, "clIcdGetPlatformIDsKHR" if hasattr(h, "clIcdGetPlatformIDsKHR") else "(no icd entry)")
    except OSError as e:
        print("direct load failed:", lib, e)
