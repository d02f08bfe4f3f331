"""Timeline of one attention CTA from the -DKOCR_TRACE build (KOCR_LIB=.../libkocr_trace.so): per warp and sub-step, the clock64
stamps at fixed points of the loop. Prints the mean duration of each phase over the steady-state sub-steps and a merged
timeline of a few sub-steps, so that the critical path of the softmax / MMA hand-shake can be read off.

softmax warps (4-11) points: 0 loop top, 1 s_full seen, 2 scores in registers, 3 row max done, 4 exponent-phase token held,
5 exponents + row sum done, 6 P stores issued, 7 o_done seen, 8 wait::st done, 9 p_full arrived.
MMA warps (1, 2) points: 0 loop top, 3 v_full seen, 4 p_full seen, 1 P.V issued, 2 next S issued."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from karanta_ocr_b200 import _lib
from tests import gpu_util as gu

pages, H, N = 4, 16, 6624
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(pages * N, H * 240, generator=g) * 0.5).to(torch.bfloat16).cuda()
cu = [i * N for i in range(pages + 1)]
for _ in range(3):
    gu.op_attention(qkv, cu, H)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
W, S, P = 16, 96, 12
buf = np.zeros((W, S, P), dtype=np.int64)
rc = lib.kocr_debug_attn_trace(C.c_void_p(buf.ctypes.data), C.c_int64(buf.nbytes))
assert rc == 0, rc
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/attn_trace.npy", buf)
lo, hi = 20, 70  # steady state
names_s = ["wait s_full (+LDTM issue in the three-tile kernel)", "LDTM+wait::ld", "mask+row max", "token wait (bar.sync)", "exponents+sum", "STTM issue", "o_done wait", "rescale+wait::st", "fence+arrive", "loop back"]
print("softmax warps: mean cycles per phase over sub-steps %d..%d" % (lo, hi))
for w in range(4, 16):
    t = buf[w]
    d = [np.mean(t[lo:hi, k + 1] - t[lo:hi, k]) for k in range(9)] + [np.mean(t[lo + 1:hi + 1, 0] - t[lo:hi, 9])]
    per = np.mean(t[lo + 1:hi + 1, 0] - t[lo:hi, 0])
    print(f"warp {w:2d} (tile {(w - 4) // 4}): period {per:7.1f} | " + " | ".join(f"{n} {x:6.1f}" for n, x in zip(names_s, d)))
print("MMA warps: mean cycles")
for w in (1, 2, 3):
    t = buf[w]
    per = np.mean(t[lo + 1:hi + 1, 0] - t[lo:hi, 0])
    print(f"warp {w}: period {per:7.1f} | wait v_full {np.mean(t[lo:hi,3]-t[lo:hi,0]):6.1f} | wait p_full {np.mean(t[lo:hi,4]-t[lo:hi,3]):6.1f} | "
          f"issue P.V {np.mean(t[lo:hi,1]-t[lo:hi,4]):6.1f} | wait k_full + issue S {np.mean(t[lo:hi,2]-t[lo:hi,1]):6.1f}")
# merged timeline of sub-steps 30..33, warps 4 and 8 (same lane quarter) + MMA warps
ev = []
for w, pts in ((5, range(10)), (9, range(10)), (13, range(10)), (1, (0, 3, 4, 1, 2)), (2, (0, 3, 4, 1, 2)), (3, (0, 3, 4, 1, 2))):
    for i in range(30, 34):
        for k in pts:
            ev.append((int(buf[w, i, k]), w, i, k))
ev.sort()
t0 = ev[0][0]
print("timeline (cycles since first event): warp, sub-step, point")
for tt, w, i, k in ev:
    print(f"{tt - t0:7d}  w{w:<2d} s{i} p{k}")
