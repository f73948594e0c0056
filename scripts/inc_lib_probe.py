"""Times coup_vec_rollout_incremental of ANY build of the library (plain ctypes, only the entry points it needs):
    python scripts/inc_lib_probe.py <path/to/libcoup_b200*.so> [steps]
Used to A/B the whole-sector variant of the incremental kernel (built from commit 866cc9a) against the shipped one under ncu."""
import ctypes as C, sys
import torch

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
lib = C.CDLL(path)


class Opts(C.Structure):
    _fields_ = [("num_envs", C.c_uint32), ("device", C.c_int32), ("seed", C.c_uint64), ("global_env_offset", C.c_uint64),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


vp = C.c_void_p
lib.coup_vec_create.argtypes = [C.POINTER(Opts), C.POINTER(vp)]
lib.coup_vec_rollout.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
lib.coup_vec_information_state_tensor_strided.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, vp]
lib.coup_vec_rollout_incremental.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, vp]
lib.coup_last_error.restype = C.c_char_p
n = 1 << 20
torch.zeros(1, device="cuda")
h = vp()
assert lib.coup_vec_create(C.byref(Opts(n, 0, 1, 0, 1, 0)), C.byref(h)) == 0
buf = torch.empty((2 * n, 2492), dtype=torch.float32, device="cuda")
ref = torch.empty_like(buf)
s = vp(torch.cuda.current_stream().cuda_stream)
assert lib.coup_vec_rollout(h, 100, -1, 0, None, s) == 0
assert lib.coup_vec_information_state_tensor_strided(h, 3, 0, vp(buf.data_ptr()), 2492, s) == 0
assert lib.coup_vec_rollout_incremental(h, 5, 0, vp(buf.data_ptr()), 2492, s) == 0, lib.coup_last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
assert lib.coup_vec_rollout_incremental(h, steps, 0, vp(buf.data_ptr()), 2492, s) == 0
e1.record()
torch.cuda.synchronize()
assert lib.coup_vec_information_state_tensor_strided(h, 3, 0, vp(ref.data_ptr()), 2492, s) == 0
torch.cuda.synchronize()
print("%s: %.3f ms/step, buffer == dense encoder: %s" % (path, e0.elapsed_time(e1) / steps, bool(torch.equal(buf, ref))))
