"""Dev tool: probe tcgen05.mma operand layouts on hardware (one MMA, M=128, K=8, tf32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctypes
from dgnn_b200._lib import ptr

# the probe kernel lives in its own library (build line: header of tools/debug_umma.cu)
_dbg = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_variants", "libdebug_umma.so"))
_dbg.dgnn_debug_umma.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                                 ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]


def call(name, *a):
    assert getattr(_dbg, name)(*a) == 0

M, K = 128, 8
DEV = "cuda:0"


def desc_bits(lbo, sbo, swz):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | swz << 61


def idesc(n, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((M >> 4) << 24)


def run(a_img, b_img, ad, bd, idc, n):
    a = torch.from_numpy(a_img.view(np.uint8)).to(DEV); b = torch.from_numpy(b_img.view(np.uint8)).to(DEV)
    out = torch.zeros(M, n, device=DEV)
    call("dgnn_debug_umma", ptr(a), a.numel(), ptr(b), b.numel(), ad, bd, idc, n, ptr(out), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def img_kmajor_sw128(X):
    """X [rows, 32 floats] -> K-major SW128 atom image (rows multiple of 8)."""
    rows = X.shape[0]
    img = np.zeros((rows * 32,), np.float32)
    for r in range(rows):
        for k in range(X.shape[1]):
            off = r * 128 + (((k >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2)
            img[off // 4] = X[r, k]
    return img


def img_mn_sw128(X, n_k=8):
    """X [k, mn] -> MN-major SW128: block b (32 mn) at b*(n_k*128); row k at k*128; chunk ^= k&7."""
    k_, mn = X.shape
    nb = (mn + 31) // 32
    img = np.zeros((nb * n_k * 32,), np.float32)
    for k in range(k_):
        for m in range(mn):
            b, mm = divmod(m, 32)
            off = b * n_k * 128 + (k >> 3) * 1024 + (k & 7) * 128 + ((((mm >> 2) ^ (k & 7))) << 4) + ((mm & 3) << 2)
            img[off // 4] = X[k, m]
    return img


def img_mn_nosw(X, lbo, sbo):
    """no swizzle MN-major: core matrix 8 k x 16 B; element (mn,k): (mn%4)*4 + (k%8)*16 + (mn//4)*sbo + (k//8)*lbo."""
    k_, mn = X.shape
    size = (mn // 4) * sbo + ((k_ + 7) // 8) * lbo + 128
    img = np.zeros((size // 4,), np.float32)
    for k in range(k_):
        for m in range(mn):
            off = (m % 4) * 4 + (k % 8) * 16 + (m // 4) * sbo + (k // 8) * lbo
            img[off // 4] = X[k, m]
    return img


rng = np.random.default_rng(0)
# tf32-exact small integers so that results are exact
A = rng.integers(-3, 4, size=(M, K)).astype(np.float32)      # D = A . B^T, A [M,K], B [N,K]
N = 32
B = rng.integers(-3, 4, size=(N, K)).astype(np.float32)
ref = A @ B.T
pad = lambda X: np.concatenate([X, np.zeros((X.shape[0], 32 - X.shape[1]), np.float32)], 1)

# 0) sanity: K-major SW128 both (the layout the layer kernels use)
d = run(img_kmajor_sw128(pad(A)), img_kmajor_sw128(pad(B)), desc_bits(16, 1024, 2), desc_bits(16, 1024, 2), idesc(N, 0, 0), N)
print("K-major/K-major SW128        max err", np.abs(d - ref).max())

# 1) A MN-major SW128 (hypothesis H1: lbo = block stride, sbo = 8-k group stride), B K-major
for lbo, sbo, name in ((1024, 1024, "H1 lbo=blk(1024) sbo=1024"), (1024, 128, "lbo=1024 sbo=128"), (128, 1024, "lbo=128 sbo=1024")):
    d = run(img_mn_sw128(A.T.copy()), img_kmajor_sw128(pad(B)), desc_bits(lbo, sbo, 2), desc_bits(16, 1024, 2), idesc(N, 1, 0), N)
    print("A MN SW128 %-28s max err %g  (nonzero %d)" % (name, np.abs(d - ref).max(), (d != 0).sum()))

# 2) B MN-major SW128, A K-major
d = run(img_kmajor_sw128(pad(A)), img_mn_sw128(B.T.copy()), desc_bits(16, 1024, 2), desc_bits(1024, 1024, 2), idesc(N, 0, 1), N)
print("B MN SW128 H1                max err", np.abs(d - ref).max(), (d != 0).sum())

# 3) both MN-major SW128
d = run(img_mn_sw128(A.T.copy()), img_mn_sw128(B.T.copy()), desc_bits(1024, 1024, 2), desc_bits(1024, 1024, 2), idesc(N, 1, 1), N)
print("both MN SW128 H1             max err", np.abs(d - ref).max(), (d != 0).sum())

# 4) no-swizzle MN-major A: core matrices 8k x 4mn (128 B), mn-core stride sbo=128, k-group stride lbo
for lbo, sbo in ((4096, 128), (128, 4096)):
    img = img_mn_nosw(A.T.copy(), lbo if lbo > sbo else 128 * 32, sbo if sbo < lbo else 128) if False else None
d = run(img_mn_nosw(A.T.copy(), 4096, 128), img_kmajor_sw128(pad(B)), desc_bits(4096, 128, 0), desc_bits(16, 1024, 2), idesc(N, 1, 0), N)
print("A MN no-swizzle lbo=4096(k) sbo=128(mn)  max err", np.abs(d - ref).max(), (d != 0).sum())
d = run(img_mn_nosw(A.T.copy(), 4096, 128), img_kmajor_sw128(pad(B)), desc_bits(128, 4096, 0), desc_bits(16, 1024, 2), idesc(N, 1, 0), N)
print("A MN no-swizzle swapped fields           max err", np.abs(d - ref).max(), (d != 0).sum())

# 5) wider: N = 256 both MN-major with 4 KB blocks (the dW kernel's stage layout, 32 cells per stage -> n_k = 32)
N2 = 256
B2 = rng.integers(-3, 4, size=(N2, K)).astype(np.float32)
ref2 = A @ B2.T
d = run(img_mn_sw128(A.T.copy(), n_k=32), img_mn_sw128(B2.T.copy(), n_k=32), desc_bits(4096, 1024, 2), desc_bits(4096, 1024, 2), idesc(N2, 1, 1), N2)
print("dW stage layout (blocks 4 KB) max err", np.abs(d - ref2).max(), (d != 0).sum())
