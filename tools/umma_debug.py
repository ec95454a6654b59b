"""Development aid: run the tcgen05 read on small cases, dump the first score tile, compare with numpy / SIMT."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle, synth, rmnet_b200
from rmnet_b200 import ops
DEV = "cuda:0"
L = rmnet_b200.lib()
L.rmnet_debug_set_umma_dump.argtypes = [ctypes.c_void_p]
L.rmnet_debug_set_umma_dump.restype = None
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)

def run(n, T, h, w, scale, prec, dump=False):
    ins = synth.memory_read_inputs(7, n, T, h, w, scale)
    mk, mv, qk, qv = ins
    dbg = torch.full((128 * 64 + 128,), float("nan"), device=DEV)
    L.rmnet_debug_set_umma_dump(dbg.data_ptr() if dump else None)
    got = ops.memory_reader_forward(*(cu(x) for x in ins), precision=prec, impl=rmnet_b200.RMNET_IMPL_UMMA)
    torch.cuda.synchronize()
    L.rmnet_debug_set_umma_dump(None)
    simt = ops.memory_reader_forward(*(cu(x) for x in ins), precision=prec, impl=rmnet_b200.RMNET_IMPL_SIMT).cpu().numpy()
    ref = oracle.memory_read(*ins, dtype=np.float64)[0]
    g = got.cpu().numpy()
    print(f"n={n} T={T} {h}x{w} scale={scale} prec={prec}: umma-vs-oracle {np.abs(g[:, :512] - ref[:, :512]).max():.3e}  "
          f"simt-vs-oracle {np.abs(simt[:, :512] - ref[:, :512]).max():.3e}  nan={int(np.isnan(g).sum())}", flush=True)
    if dump:
        N = h * w
        S = dbg[:128 * 64].view(128, 64).cpu().numpy()
        Q = qk[0].reshape(128, N)[:, :128].T.astype(np.float64)      # [q, c]
        K = mk[0].reshape(128, T * N)[:, :64].T.astype(np.float64)   # [m, c]
        Sref = Q @ K.T
        nq = min(128, N)
        d = np.abs(S[:nq] - Sref[:nq])
        print("  S dump: max|S-Sref| =", d.max(), " |Sref|max =", np.abs(Sref).max(), " nan in S:", int(np.isnan(S[:nq]).sum()))
        if d.max() > 1e-2:
            # try to recognise a permutation: which reference column best matches dumped column j?
            for j in (0, 1, 2, 8, 16, 33):
                best = np.argmin(np.abs(Sref[:nq] - S[:nq, j:j + 1]).sum(0))
                print(f"   dumped col {j} ~ ref col {best} (err {np.abs(Sref[:nq, best] - S[:nq, j]).max():.3e})")
            for i in (0, 1, 2, 8, 33, 64):
                best = np.argmin(np.abs(Sref[:nq] - S[i:i + 1]).sum(1))
                print(f"   dumped row {i} ~ ref row {best} (err {np.abs(Sref[best] - S[i]).max():.3e})")
    return g, ref

run(1, 1, 8, 16, 0.3, 0, dump=True)     # N = 128, M = 128: one q tile, two KV tiles
run(1, 1, 8, 16, 0.3, 1, dump=True)
run(1, 1, 4, 8, 0.3, 0)                 # ragged: N = 32
run(2, 3, 15, 27, 0.5, 0)               # config-1 shape
run(3, 5, 30, 54, 0.5, 0)               # config-2 shape
run(3, 5, 30, 54, 1.0, 0)
run(1, 2, 30, 54, 3.0, 0)               # huge scores: exercises the lazy rescale
