"""single-window dump of the tcgen05 window-attention kernel vs torch (development aid)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from occformer_b200 import ops
from occformer_b200._lib import lib
torch.manual_seed(0)
dev = torch.device("cuda:0")
B, X, Y, Z, C = 1, 7, 7, 1, 32
rows = B * X * Y * (Z + 1)
qkv = torch.randn(rows, 3 * C, device=dev)
qb = torch.randn(3 * C, device=dev)
bias = torch.zeros(1, 2404, device=dev)
dbg = torch.full((128, 192), -7.0, device=dev)
lib().occ_window_attention_set_debug(ctypes.c_void_p(dbg.data_ptr()))
out = ops.window_attention(qkv, qb, bias, B, X, Y, Z, C, 1, False)
torch.cuda.synchronize()
lib().occ_window_attention_set_debug(None)
d = dbg.cpu()
q, k, v = qkv[:, :C].cpu(), qkv[:, C:2 * C].cpu(), qkv[:, 2 * C:].cpu()
# window A = image z=0: token rows r = ((x*Y)+y)*Z + 0 = t ; window B = BEV rows 49 + t
SA = q[:49] @ k[:49].t()
print("S raw win A: kernel vs ref (row0 first 6):", d[0, :6].tolist(), SA[0, :6].tolist())
print("max |S-ref| win A:", float((d[:49, :49] - SA).abs().max()))
SB = q[49:98] @ k[49:98].t()
print("max |S-ref| win B:", float((d[64:113, :49] - SB).abs().max()))
PA = torch.softmax(SA * 32 ** -0.5, -1)
print("P row0 kernel (unnormalised)/sum vs ref:", (d[0, 64:70] / d[0, 160]).tolist(), PA[0, :6].tolist())
print("sum,m,row,reg row0:", d[0, 160:164].tolist(), " row 64:", d[64, 160:164].tolist())
OA = PA @ v[:49]
print("O raw row0 / sum:", (d[0, 128:134] / d[0, 160]).tolist(), " ref:", OA[0, :6].tolist())
print("out row0:", out[0, :6].tolist())
print("max |out - ref| A:", float((out[:49].cpu() - OA).abs().max()))
