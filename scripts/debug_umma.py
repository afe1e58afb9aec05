import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occformer_b200._lib import lib
torch.manual_seed(0)
dev = torch.device("cuda:0")
def r(t):
    i = t.clone().view(torch.int32); i.add_(0x1000).bitwise_and_(-8192); return i.view(torch.float32)
A = r(torch.randn(128, 64)).to(dev); V = r(torch.randn(64, 32)).to(dev)
ref = (A.double() @ V.double()).float()
for mode in range(8):
    D = torch.full((128, 32), -7.0, device=dev)
    rc = lib().occ_debug_umma_probe(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(V.data_ptr()), ctypes.c_void_p(D.data_ptr()), mode, None)
    torch.cuda.synchronize()
    err = float((D - ref).abs().max())
    print(f"mode {mode}: rc={rc} max|D-ref|={err:.3e}  D[0,:4]={D[0,:4].tolist()} ref={ref[0,:4].tolist()}")
    if err > 1e-2:
        # is it a permutation / transpose of the right answer?
        print("   |D| mean", float(D.abs().mean()), " ref mean", float(ref.abs().mean()))
