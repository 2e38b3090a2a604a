"""Per-shape device time of the tcgen05 GEMM (matrix and implicit 3x3 conv A-operand) on the shapes that dominate one C2
step, next to torch.matmul (cuBLAS) on the same operands. Tile policy is chosen through the library's own environment
switches (SDB_GEMM_BN, SDB_GEMM_DEEP, SDB_GEMM_BN256, ...), so run one process per policy. Diagnostic only.
    python tools/gemm_shapes.py [tag]"""
import csv, ctypes as C, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import lib as L, nn_ops as O

dev = torch.device("cuda:0")
lib = L.load()
# (kind, M or (n,h,w), N, K or Cin, launches per step)
SHAPES = [
    ("mm", 8192, 2560, 2880, 0),
    ("mm", 20480, 320, 320, 25), ("mm", 5120, 640, 640, 25), ("mm", 1280, 1280, 1280, 25),
    ("mm", 20480, 2560, 320, 5), ("mm", 5120, 5120, 640, 5), ("mm", 1280, 10240, 1280, 5),
    ("mm", 20480, 320, 1280, 5), ("mm", 5120, 640, 2560, 5), ("mm", 1280, 1280, 5120, 5),
    ("mm", 20480, 960, 320, 5), ("mm", 5120, 1920, 640, 5), ("mm", 1280, 3840, 1280, 5),
    ("mm", 385, 2560, 1024, 6),
    ("conv", (1, 512, 512), 128, 128, 8), ("conv", (1, 256, 256), 256, 256, 6), ("conv", (1, 128, 128), 512, 512, 6),
    ("conv", (1, 64, 64), 512, 512, 16),
    ("conv", (5, 64, 64), 320, 320, 7), ("conv", (5, 32, 32), 640, 640, 6), ("conv", (5, 16, 16), 1280, 1280, 7),
    ("conv", (5, 8, 8), 1280, 1280, 11), ("conv", (5, 16, 16), 1280, 2560, 2), ("conv", (5, 32, 32), 640, 1920, 1),
    ("conv", (5, 64, 64), 320, 960, 1),
]
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
rows = []
tot = 0.0
for kind, m, N, K, per_step in SHAPES:
    if kind == "mm":
        a = torch.randn(m, K, device=dev, dtype=torch.float16) * 0.1
        b = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.1
        bias = torch.randn(N, device=dev, dtype=torch.float16)
        out = torch.empty(m, N, device=dev, dtype=torch.float16)
        run = lambda: O.gemm(a, b, bias=bias, out=out)
        M, Kt = m, K
        ref = lambda: torch.matmul(a, b.t())
    else:
        n, h, w = m
        x = torch.randn(n, h, w, K, device=dev, dtype=torch.float16) * 0.1
        wt = torch.randn(N, 3, 3, K, device=dev, dtype=torch.float16) * 0.1
        bias = torch.randn(N, device=dev, dtype=torch.float16)
        run = lambda: O.conv3x3(x, wt, bias=bias)
        M, Kt = n * h * w, 9 * K
        a2 = torch.randn(M, Kt, device=dev, dtype=torch.float16) * 0.1
        b2 = wt.reshape(N, Kt)
        ref = lambda: torch.matmul(a2, b2.t())
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    path = tempfile.mktemp(suffix=".csv")
    lib.sdb_gemm_profile_begin()
    lib.sdb_gemm_profile_dump(path.encode())
    for _ in range(15):
        run()
    gm, gf, gl = C.c_double(), C.c_double(), C.c_int()
    L.check(lib.sdb_gemm_profile_end(C.byref(gm), C.byref(gf), C.byref(gl)), "profile_end")
    r = list(csv.DictReader(open(path)))
    os.unlink(path)
    ms = sorted(float(x["ms"]) for x in r)[len(r) // 2]
    for _ in range(3):
        ref()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ref()
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / 20
    fl = 2.0 * M * N * Kt
    tot += ms * per_step
    rows.append((kind, M, N, Kt, r[0]["bn"], r[0]["splits"], ms * 1e3, fl / ms / 1e9, ms_ref * 1e3, fl / ms_ref / 1e9, per_step))
    print(f"{tag} {kind:4s} M{M:<7d} N{N:<6d} K{Kt:<6d} bn{r[0]['bn']:>4s} sp{r[0]['splits']:>2s}  {ms*1e3:8.1f} us {fl/ms/1e9:7.0f} TF | cublas {ms_ref*1e3:8.1f} us {fl/ms_ref/1e9:7.0f} TF | x{per_step}", flush=True)
print(f"{tag} weighted ms/step over these shapes: {tot:.3f}")
