import ctypes, glob, os, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
path = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = ctypes.CDLL(path[0] if path else "libcudart.so")
v = ctypes.c_size_t(0)
print("get before:", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value)
for g in (32, 64, 128):
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(g))
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("set", g, "rc", rc, "now", v.value)
