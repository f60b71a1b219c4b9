"""Plain-PyTorch restatement of the reference's optimizer step (test infrastructure): global-norm clip
(torch.nn.utils.clip_grad_norm_, train_concap.py:305-307) + pytorch_transformers.AdamW (train_concap.py:227): bias-corrected
Adam, decoupled weight decay applied to the updated parameter, no decay on biases / LayerNorm unless told otherwise.
pytorch_transformers 1.1.0 is not in the image (requirements.txt:39): restated from its published algorithm."""
import torch

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")


def adamw_step(named, state, lr, step, clip=5.0, betas=(0.9, 0.999), eps=1e-6, wd=0.01, lr_of=None, wd_of=None):
    grads = [p.grad for _, p in named if p.grad is not None]
    scale = 1.0
    if clip and clip > 0:
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).item()
        scale = min(1.0, clip / (total + 1e-6))
    for name, p in named:
        if p.grad is None:
            continue
        g = p.grad * scale
        m, v = state.setdefault(name, (torch.zeros_like(p), torch.zeros_like(p)))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        plr = lr_of(name) if lr_of else lr
        step_size = plr * (1 - betas[1] ** step) ** 0.5 / (1 - betas[0] ** step)
        decay = wd_of(name) if wd_of else (0.0 if any(nd in name for nd in NO_DECAY) else wd)
        with torch.no_grad():
            p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)
            if decay > 0:
                p.add_(p, alpha=-plr * decay)
