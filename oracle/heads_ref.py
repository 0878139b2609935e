"""fp32 torch restatement of the reference's contrastive heads and DRO losses.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines it follows; quirks are reproduced, not fixed.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- A.1
def pair_nll(q, a, b):
    """ANCE/model/models.py:101-108.  Returns (loss[B], accs[B] int64, logits[B,2])."""
    logits = torch.stack([(q * a).sum(-1), (q * b).sum(-1)], dim=1)
    loss = -F.log_softmax(logits, dim=1)[:, 0]
    accs = torch.argmax(logits, dim=1)
    return loss, accs, logits


def erm_reduce(loss, weights):
    """ANCE/model/models.py:259-262: mean(loss * weights)."""
    return (loss * weights).mean()


# --------------------------------------------------------------------------- A.2
def coco_target(n):
    """COCO/modeling.py:172-173: [1,0,3,2,...]."""
    return torch.arange(n, dtype=torch.long).view(-1, 2).flip([1]).flatten().contiguous()


def coco_contrastive(E, world_size=1):
    """COCO/modeling.py:244-248.  E is the (gathered) [N,H] matrix; returns loss[N]."""
    S = E @ E.t()
    S = S.clone()
    S.fill_diagonal_(float("-inf"))
    return F.cross_entropy(S, coco_target(E.shape[0]), reduction="none") * world_size


# --------------------------------------------------------------------------- A.3
def qp_infonce(Q, P, targets=None):
    """q x p in-batch InfoNCE (north-star K9'; not in the reference -> torch restatement)."""
    if targets is None:
        targets = torch.arange(Q.shape[0])
    return F.cross_entropy(Q @ P.t(), targets, reduction="none")


# --------------------------------------------------------------------------- A.4
def group_stats(losses, g, n_groups):
    """dro_loss.py:217-224: scatter_add sums / counts / means with 0-count guard."""
    zero = losses.new_zeros(n_groups)
    sums = zero.scatter_add(0, g, losses)
    cnts = zero.scatter_add(0, g, torch.ones_like(losses))
    means = sums / (cnts + (cnts == 0).float())
    return sums, cnts, means


def idro_weight_update(h_fun, group_grads, means, cnts, alpha, ema, rho, eps):
    """dro_loss.py:235-251 given the (all-reduced) [G, P] per-group gradient matrix."""
    norm = torch.linalg.norm(group_grads, dim=-1, keepdim=True)
    gh = group_grads / (1e-12 + norm)
    rtg = gh @ gh.t()
    gl = torch.pow(means.detach().unsqueeze(-1), alpha)
    rtg = (gl @ gl.t()) * rtg
    e = rho * rtg.mean(dim=0)
    e = e * (cnts > 0).float()
    e = e - e.max()
    w = torch.exp(e)
    h = torch.pow(h_fun, ema) * w * (cnts != 0).float()
    h = h / h.sum()
    return torch.clamp(h, min=eps)


def idro_forward(losses, g, params, h_fun, n_groups, alpha, ema, rho, eps, all_reduce=None):
    """dro_loss.py:216-254 (training mode).

    ``losses`` must carry an autograd graph down to ``params`` (the "last
    layers" selected by iDROLoss._params, dro_loss.py:174-190).  Returns
    (robust_loss, means, cnts, new_h_fun).  ``all_reduce`` is an optional
    callable applied to the [G,P] matrix (dist.all_reduce at :232).
    """
    sums, cnts, means = group_stats(losses, g, n_groups)
    robust = (means * h_fun).sum()
    rows = []
    dim = sum(p.numel() for p in params)
    for gi in range(n_groups):
        if cnts[gi] > 0:
            gr = torch.autograd.grad(means[gi], params, retain_graph=True)
            rows.append(torch.cat([x.reshape(-1) for x in gr]))
        else:
            rows.append(torch.zeros(dim))
    G = torch.stack(rows).detach()
    if all_reduce is not None:
        G = all_reduce(G)
    new_h = idro_weight_update(h_fun, G, means, cnts, alpha, ema, rho, eps)
    return robust, means.detach(), cnts.detach(), new_h


def idro_param_names(names, model_size="base"):
    """dro_loss.py:174-190: substring match on 'layer.N' (note: 'layer.1' style
    collisions do not occur for 9/10/11 or 22/23)."""
    select = ["layer.23", "layer.22"] if model_size == "large" else ["layer.10", "layer.11", "layer.9"]
    return [n for n in names if any(n.find(s) >= 0 for s in select)]


# --------------------------------------------------------------------------- A.5
def dro_greedy_forward(losses, g, h_fun, sum_losses, count_cat, n_groups, alpha, eps, ema,
                       weight_ema=True, w=None, gathered=None):
    """dro_loss.py:49-120 (training mode).  ``gathered`` = (g_all, losses_all) if the
    caller emulates the all_gather at :64-65; default is world_size 1.
    Returns (robust, group_losses, group_counts, new_h, new_sum_losses, new_count_cat)."""
    if w is not None:
        losses = losses * w
    B = losses.shape[0]
    zero = losses.new_zeros(n_groups)
    gsum = zero.scatter_add(0, g, losses)
    robust = (gsum * h_fun).sum() / B
    with torch.no_grad():
        g_all, l_all = gathered if gathered is not None else (g, losses.detach())
        cnt_agg = zero.scatter_add(0, g_all, torch.ones_like(l_all))
        los_agg = zero.scatter_add(0, g_all, l_all)
        gl = los_agg / (cnt_agg + (cnt_agg == 0).float())
        valid = cnt_agg > 0
        sum_losses = sum_losses.clone()
        sum_losses[valid] = sum_losses[valid] * (1 - ema) + gl[valid] * ema
        count_cat = count_cat * (1 - ema) + cnt_agg * ema
        # update_mw (:88-120)
        frac = count_cat / count_cat.sum()
        sl, sid = torch.sort(sum_losses, descending=True)
        sf = frac[sid]
        cutoff = int(torch.sum(torch.cumsum(sf, 0) < alpha))
        if cutoff == len(sf):
            cutoff = len(sf) - 1
        tmp = torch.full_like(h_fun, eps)
        tmp[sid[:cutoff]] = 1.0 / alpha
        leftover = 1.0 - sf[:cutoff].sum() / alpha
        tmp[sid[cutoff]] = max(float(leftover / sf[cutoff]), eps)
        if weight_ema:
            tmp = tmp.clamp(min=eps)
            new_h = h_fun * (1 - ema) + tmp * ema
        else:
            new_h = tmp
        cnts = zero.scatter_add(0, g, torch.ones_like(losses))
        gl_local = gsum.detach() / (cnts + (cnts == 0).float())
    return robust, gl_local, cnts, new_h, sum_losses, count_cat
