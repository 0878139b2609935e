"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): fp32 torch restatement of the optimizers on the training path.

* ``lamb_step``      follows ANCE/utils/lamb.py:71-121 line by line (no bias correction :101-103, weight norm
                     clamped to [0, 10] :105, adam_step = m / (sqrt(v) + eps) (+ wd * p) :107-109, trust ratio 1 when
                     either norm is 0 :112-115, p -= lr * trust * adam_step :121).
* ``hf_adamw_step``  follows transformers.AdamW (the ``AdamW`` the reference imports, ANCE/drivers/run_ann.py:19,
                     139-144; transformers==2.3.0 optimization.py): m, v EMA; step = lr * sqrt(bc2) / bc1;
                     p -= step * m / (sqrt(v) + eps); then p -= lr * wd * p.
* ``clip_coef``      torch.nn.utils.clip_grad_norm_ (run_ann.py:345-352): min(1, max_norm / (total_norm + 1e-6)).

Pinning: ``lamb_step`` is checked against the UNMODIFIED reference class (ANCE/utils/lamb.py runs here once the absent
tensorboardX import is stubbed; its deprecated ``add_(Number, Tensor)`` overloads still work): fixture
tests/golden/lamb_tiny.npz written by oracle/make_golden.py, test tests/test_oracle_golden.py.  The torch-semantics
AdamW is checked against the installed ``torch.optim.AdamW`` (tests/test_optim_gpu.py).  ``hf_adamw_step`` stays
unpinned: transformers 5.x no longer ships AdamW (restated from transformers==2.3.0 optimization.py).
"""
import math

import torch


def lamb_step(p, g, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0):
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    weight_norm = p.pow(2).sum().sqrt().clamp(0, 10)
    adam_step = m / v.sqrt().add(eps)
    if weight_decay != 0:
        adam_step.add_(p, alpha=weight_decay)
    adam_norm = adam_step.pow(2).sum().sqrt()
    trust = 1.0 if (weight_norm == 0 or adam_norm == 0) else float(weight_norm / adam_norm)
    p.add_(adam_step, alpha=-lr * trust)
    return trust


def hf_adamw_step(p, g, m, v, step, *, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0):
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0:
        p.add_(p, alpha=-lr * weight_decay)


def clip_coef(grads, max_norm):
    total = torch.sqrt(sum((g.float() ** 2).sum() for g in grads))
    return min(1.0, max_norm / (float(total) + 1e-6)), float(total)
