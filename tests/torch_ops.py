"""Plain-torch stand-in for ``cfpnet_b200.train.CudaOps`` (TEST INFRASTRUCTURE: same method names and semantics, float64 on
the CPU) so that the op SEQUENCING of ``cfpnet_b200/train_seq.py`` can be held to the reference's ``.train()`` gradients
without a GPU.  The product never imports this."""
import torch
import torch.nn.functional as F

from oracle import cfp_oracle as O
from oracle import cfp_oracle_bwd as OB


class TorchOps:
    def linear(self, x, w, bias=None, acc=None):
        y = x @ w.t()
        if bias is not None:
            y = y + bias
        if acc is not None:
            acc += y
            return acc
        return y

    def linear_dx(self, dy, w):
        return dy @ w

    def linear_dw(self, dy, x):
        return dy.t() @ x

    def colsum(self, x):
        return x.sum(0)

    def ln_fwd(self, x, g, b, eps):
        return F.layer_norm(x, (x.shape[-1],), g, b, eps)

    def ln_bwd(self, x, g, dy, eps):
        return OB.layer_norm_bwd(x, g, dy, eps)

    def bn_fwd(self, x, bn, relu):
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        rstd = torch.rsqrt(var + bn.eps)
        y = (x - mean) * rstd * bn.weight.detach() + bn.bias.detach()
        return (torch.relu(y) if relu else y), mean, rstd

    def bn_bwd(self, dy, x, mean, rstd, bn, relu):
        xh = (x - mean) * rstd
        g = dy
        if relu:
            g = g * ((xh * bn.weight.detach() + bn.bias.detach()) > 0)
        n = x.shape[0]
        s1, s2 = g.sum(0), (g * xh).sum(0)
        return bn.weight.detach() * rstd / n * (n * g - s1 - xh * s2), s2, s1

    def ew(self, a, b, op):
        if op == "add":
            return a + b
        if op == "relu":
            return torch.relu(a)
        if op == "relu_mask":
            return a * (b > 0)
        if op == "elu1":
            return F.elu(a) + 1
        if op == "elu1_grad_mul":
            return a * OB.elu1_grad(b)
        if op == "neg_div":
            return -a / b
        raise KeyError(op)

    def attn_reduce(self, A, Bm, w, G, R, nh):
        C = A.shape[1]
        d = C // nh
        Ah, Bh = A.view(G, R, nh, d), Bm.view(G, R, nh, d)
        KV = torch.einsum("grhi,grhj->ghij", Ah, Bh)
        if w is None:
            As = A.view(G, R, C).sum(1)
        else:
            As = torch.einsum("grh,grhi->ghi", w.view(G, R, nh), Ah).reshape(G, C)
        return KV, As

    def attn_apply(self, X, KV, G, R, nh, transpose):
        C = X.shape[1]
        d = C // nh
        Xh = X.view(G, R, nh, d)
        out = torch.einsum("grhj,ghij->grhi", Xh, KV) if transpose else torch.einsum("grhi,ghij->grhj", Xh, KV)
        return out.reshape(G * R, C)

    def head_dot(self, a, b, nh, rows_per_group, eps):
        C = a.shape[1]
        d = C // nh
        if rows_per_group:
            b = b.repeat_interleave(rows_per_group, dim=0)
        return (a.view(-1, nh, d) * b.view(-1, nh, d)).sum(-1) + eps

    def head_scale(self, a, s, nh, divide):
        d = a.shape[1] // nh
        s = s.repeat_interleave(d, dim=1)
        return a / s if divide else a * s

    def head_axpy(self, out, s, b, nh, rows_per_group):
        d = out.shape[1] // nh
        out += s.repeat_interleave(d, dim=1) * b.repeat_interleave(rows_per_group, dim=0)

    def group_add(self, out, b, rows_per_group):
        out += b.repeat_interleave(rows_per_group, dim=0)

    def group_scale(self, x, m, rows_per_group):
        return x * m.repeat_interleave(rows_per_group).unsqueeze(1)

    def gather_rows(self, src, idx):
        idx = idx.long()
        out = src[idx.clamp_min(0)]
        return out * (idx >= 0).unsqueeze(1)

    def scatter_add_rows(self, src, idx, base):
        idx = idx.long()
        ok = idx >= 0
        out = base.clone()
        out.index_add_(0, idx[ok], src[ok])
        return out

    def zeros_like(self, t):
        return torch.zeros_like(t)

    def index_mod(self, n, S, device):
        return (torch.arange(n) % S).to(torch.int32)

    def posenc_tokens(self, x, pos, max_res, oy, ox):
        B, C, H, W = x.shape
        p = pos.view(max_res[0], max_res[1], C)[oy:oy + H, ox:ox + W]
        return (x.permute(0, 2, 3, 1) + p).reshape(B * H * W, C)

    def nchw_to_tokens(self, x):
        B, C, H, W = x.shape
        return x.permute(0, 2, 3, 1).reshape(B * H * W, C)

    def tokens_to_nchw(self, t, B, C, H, W):
        return t.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()

    # LKPM: the oracle's restatements (the CUDA side has its own, GPU-tested sequencing: cfpnet_b200/train.py)
    def lkpm_fwd(self, blk, x_tok, B, H, W):
        p = {k: v.detach() for k, v in blk.state_dict().items()}
        N = H * W
        out = O.lkpm(p, x_tok.view(B, N, -1), H, W, bn_stats={})
        return out.reshape(B * N, -1), (p, x_tok)

    def lkpm_bwd(self, blk, saved, d, B, H, W):
        p, x_tok = saved
        N = H * W
        dx, g = OB.lkpm_bwd(p, x_tok.view(B, N, -1), H, W, d.view(B, N, -1))
        return dx.reshape(B * N, -1), g
