"""Dev tool: cycles per tcgen05.mma when one (or several) threads issue a stream of them (tools/bench_umma.cu)."""
import ctypes, os, sys
import torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, "gpurun_variants", "libbench_umma.so"))
lib.bench_umma.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
out = torch.zeros(148, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
names = {0: "SW128", 1: "SW32", 2: "SW64", 3: "noswz"}
for grid in (148,):
    for bf16 in (0, 1):
        for a_tmem in (0, 1):
            for nw_log in (0, 1, 2):
                for n in (32, 64, 112):
                    v = 0 | (bf16 << 2) | (a_tmem << 3) | (nw_log << 6)
                    for _ in range(2):
                        assert lib.bench_umma(v, n, 4000, out.data_ptr(), grid, st) == 0
                        torch.cuda.synchronize()
                    c = out[:grid].mean().item()
                    print("grid %3d %s A-%s issuing warps=%d N=%3d : %7.1f cycles / MMA (aggregate)" % (
                        grid, "bf16(K16)" if bf16 else "tf32(K8) ", "tmem" if a_tmem else "smem", 1 << nw_log, n, c))
