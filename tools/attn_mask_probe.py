"""Debug probe: recover the effective dropout mask of cdr_attn_fwd (K = 0 -> uniform softmax, V = one-hot) and compare
it element by element with oracle/dropout_ref.attention_mask."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cocodr_b200 import kernels as k
from oracle import dropout_ref

n_seq, L, heads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
H, T = heads * 64, n_seq * L
SEED, OFF, P, site = 20240611, 5, 0.1, 13
state = torch.tensor([SEED, OFF], dtype=torch.int64, device="cuda")
ref = dropout_ref.attention_mask(n_seq, heads, L, site, SEED, OFF, P)
got = torch.zeros(n_seq, heads, L, L)
for blk in range((L + 63) // 64):
    qkv = torch.zeros(n_seq, L, 3, heads, 64)
    for j in range(64):
        if blk * 64 + j < L:
            qkv[:, blk * 64 + j, 2, :, j] = 1.0
    qkv = qkv.reshape(T, 3 * H).half().cuda()
    out = torch.zeros(T, H, dtype=torch.float16, device="cuda")
    lse = torch.zeros(n_seq, heads, L, device="cuda")
    k.attn_fwd(qkv, None, out, lse, n_seq=n_seq, seq_len=L, heads=heads, drop=k.drop_args(state, site, P),
               drop_bits=k.attn_dropout_bits(n_seq, heads, L, "cuda"))
    o = out.float().cpu().view(n_seq, L, heads, 64).permute(0, 2, 1, 3) * L  # [n_seq, heads, L, 64] ~ mask * scale
    w = min(64, L - blk * 64)
    got[:, :, :, blk * 64:blk * 64 + w] = o[:, :, :, :w]
gm, rm = got > 0.5, ref > 0.5
print("keep frac got/ref", gm.float().mean().item(), rm.float().mean().item(), "mismatch", (gm != rm).float().mean().item())
bad = (gm != rm).nonzero()
print("first mismatches", bad[:20].tolist())
if len(bad):
    print("by column%8", torch.bincount(bad[:, 3] % 8, minlength=8).tolist(), "by col//8", torch.bincount(bad[:, 3] // 8, minlength=L // 8).tolist())
    print("by row%32", torch.bincount(bad[:, 2] % 32, minlength=32).tolist())
