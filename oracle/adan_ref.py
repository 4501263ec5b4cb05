"""oracle/adan_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the parameter update that follows the hot path: nn.utils.clip_grad_norm_(max_norm) (nerf/utils.py:984) followed by
Adan.step() with foreach=False (optimizer.py:102-198 -> _single_tensor_adan :201-256), on torch fp32 tensors.  Pinned against
tests/golden/adan.npz, which tests/golden/make_golden_adan.py records from the reference's own optimizer.py."""
import math

import torch


class AdanRef:
    def __init__(self, groups, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0, no_prox=False, clip_grad_norm=0.0):
        """groups: list of dicts {'params': [tensors], 'lr': float}"""
        self.groups = [dict(g, step=0) for g in groups]
        self.betas, self.eps, self.wd, self.max_grad_norm, self.no_prox, self.clip_grad_norm = betas, eps, weight_decay, max_grad_norm, no_prox, clip_grad_norm
        self.state = {}

    @torch.no_grad()
    def step(self, grads):
        """grads: list (per group) of lists of tensors; they are modified in place like p.grad is in the reference"""
        flat = [g for gg in grads for g in gg]
        if self.clip_grad_norm > 0:                                     # torch.nn.utils.clip_grad_norm_
            total = torch.sqrt(sum((g.float() ** 2).sum() for g in flat))
            coef = torch.clamp(self.clip_grad_norm / (total + 1e-6), max=1.0)
            for g in flat:
                g.mul_(coef)
        if self.max_grad_norm > 0:                                      # optimizer.py:112-129
            gn = torch.zeros(1)
            for g in flat:
                gn.add_(g.pow(2).sum())
            gn = torch.sqrt(gn)
            clip = float(torch.clamp(torch.tensor(self.max_grad_norm) / (gn + self.eps), max=1.0))
        else:
            clip = 1.0
        b1, b2, b3 = self.betas
        for group, gg in zip(self.groups, grads):
            group["step"] += 1
            bc1, bc2, bc3s = 1.0 - b1 ** group["step"], 1.0 - b2 ** group["step"], math.sqrt(1.0 - b3 ** group["step"])
            lr = group["lr"]
            for p, g in zip(group["params"], gg):
                st = self.state.setdefault(id(p), {})
                if not st:
                    st.update(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p), exp_avg_diff=torch.zeros_like(p))
                if "neg_pre_grad" not in st or group["step"] == 1:
                    st["neg_pre_grad"] = g.clone().mul_(-clip)
                m, v, d, n = st["exp_avg"], st["exp_avg_sq"], st["exp_avg_diff"], st["neg_pre_grad"]
                g.mul_(clip)
                n.add_(g)
                m.mul_(b1).add_(g, alpha=1 - b1)
                d.mul_(b2).add_(n, alpha=1 - b2)
                n.mul_(b2).add_(g)
                v.mul_(b3).addcmul_(n, n, value=1 - b3)
                denom = (v.sqrt() / bc3s).add_(self.eps)
                if self.no_prox:
                    p.mul_(1 - lr * self.wd)
                    p.addcdiv_(m, denom, value=-lr / bc1)
                    p.addcdiv_(d, denom, value=-lr * b2 / bc2)
                else:
                    p.addcdiv_(m, denom, value=-lr / bc1)
                    p.addcdiv_(d, denom, value=-lr * b2 / bc2)
                    p.div_(1 + lr * self.wd)
                n.zero_().add_(g, alpha=-1.0)
