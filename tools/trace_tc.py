"""Timeline of one CTA of conv_taps_tc_kernel (clock64 at pipeline events); PCODEC_TC_DEBUG bit 6."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from progressivecodec_b200 import _lib as L
from progressivecodec_b200.engine import Engine, Act, new_act, pack_conv2d

dev = torch.device("cuda", 0)
cin, cout, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dbg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
split = int(sys.argv[5]) if len(sys.argv) > 5 else 3
B = 37
m = nn.Conv2d(cin, cout, k, 1, k // 2)
pc = pack_conv2d(m, dev, "x").attach_tc(split)
x = Act(torch.randn(B, 32, 48, cin, device=dev))
out = new_act(B, 32, 48, cout, dev)
E = Engine(dev, 2)
os.environ["PCODEC_TC_VERBOSE"] = "1"
E.conv(pc, [x], out, L.EPI_GELU)
del os.environ["PCODEC_TC_VERBOSE"]
torch.cuda.synchronize()
os.environ["PCODEC_TC_DEBUG"] = str(dbg | 64)
E.conv(pc, [x], out, L.EPI_GELU)
torch.cuda.synchronize()
lib = L.lib()._lib if hasattr(L.lib(), "_lib") else L.lib()
fn = lib.pcodec_debug_tc_trace
fn.restype = C.c_int
n = fn(None, 0)
buf = (C.c_longlong * n)()
fn(buf, n)
NS, NE = 128, 12
g = [buf[NS * NE + i] for i in range(5)]
t0 = g[0]
print(f"cin {cin} cout {cout} k{k} dbg {dbg} split {split}")
print("global: setup_done %d  tmem_full %d  epilogue_end %d  exit %d" % tuple(v - t0 for v in g[1:5]))
names = ["ld_empty", "ld_pub", "cv_top", "cv_rawfull", "cv_aempty", "cv_stdone", "cv_arrived", "mma_top", "mma_afull", "mma_bfull", "mma_commit", "mma_mmas"]
print("epilogue (rel. tmem_full): setup %d | chunk0: tmem-ld %d  sts %d  pass0[lds %d  math %d  stg %d]  chunk-end %d" % tuple(buf[NS*NE+i] - g[2] for i in (9,5,6,10,11,7,8)))
print("slab 9 per-MMA issue times (rel. to mma_bfull):", [buf[NS * NE + 5 + i] - buf[9 * NE + 9] for i in range(12)])
print("slab " + " ".join(f"{n:>10s}" for n in names))
slabs = k * k * ((cin + 31) // 32)
prev = None
for s in range(min(slabs, 40)):
    row = [buf[s * NE + e] - t0 for e in range(12)]
    print(f"{s:4d} " + " ".join(f"{v:10d}" for v in row))
