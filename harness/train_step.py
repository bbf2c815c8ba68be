"""One stage-3 training iteration as the reference runs it (command_handlers/train_or_resume.py:404-460): forward,
``CE + w1*MSE + w2*FA``, backward, SGD step -- with the FA loss pluggable, so the B200 drop-in and a plain-PyTorch
restatement can be compared on identical weights and inputs.  Synthetic data of BASELINE configs[4]'s shapes."""
import torch
import torch.nn as nn

from .dsrl_model import DSRL

IGNORE = 255                    # datasets/Cityscapes/settings.py IGNORE_CLASS_LABEL
W1, W2 = 0.1, 1.0               # settings.py DEFAULT_LOSS_WEIGHTS
LR, MOMENTUM, WEIGHT_DECAY = 0.01, 0.9, 0.0005      # settings.py:38-41


def synthetic_batch(batch, device, seed, in_hw=(256, 512), num_classes=19):
    """input_image (B,3,256,512), input_org (B,3,512,1024), target (B,512,1024) uint8 with 10 % ignore."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    h, w = in_hw
    img = torch.randn((batch, 3, h, w), device=device, generator=g)
    org = torch.randn((batch, 3, 2 * h, 2 * w), device=device, generator=g)
    target = torch.randint(0, num_classes, (batch, 2 * h, 2 * w), device=device, generator=g, dtype=torch.uint8)
    target.masked_fill_(torch.rand((batch, 2 * h, 2 * w), device=device, generator=g) < 0.1, IGNORE)
    return img, org, target


class Stage3Step:
    def __init__(self, fa_loss, device, seed=54321, num_classes=19, ddp=False, ce_loss=None, fused_losses=False):
        torch.manual_seed(seed)                                   # all ranks build the same weights (train_or_resume.py:31)
        self.model = DSRL(3, num_classes).to(device).train()
        self.core = self.model
        if ddp:
            self.model = nn.parallel.DistributedDataParallel(self.model, device_ids=[device.index])
        # ce_loss: a replacement for the reference's t.nn.CrossEntropyLoss(ignore_index=...) (train_or_resume.py:116), e.g. the
        # one-pass dsrl-b200 CrossEntropyLoss, which also takes the uint8 target as is (no .long() copy)
        self.ce = ce_loss if ce_loss is not None else nn.CrossEntropyLoss(ignore_index=IGNORE)
        self.ce_takes_uint8 = ce_loss is not None
        self.mse = nn.MSELoss()
        self.fa = fa_loss
        # fused_losses: CE, MSE, both feature transformers and FA through dsrl-b200's Stage3Loss (shared passes)
        self.stage3 = None
        if fused_losses:
            from dualsuperreslearningforsemseg_b200.models.losses import Stage3Loss
            self.stage3 = Stage3Loss(self.core.SSSR_feature_transformer, self.core.SISR_feature_transformer, ignore_index=IGNORE)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=LR, momentum=MOMENTUM, weight_decay=WEIGHT_DECAY)

    def losses(self, img, org, target):
        if self.stage3 is not None:
            sssr, sisr, _, _ = self.model(img, apply_transformers=False)
            ce, mse, fa = self.stage3(sssr, sisr, target, org)
            return ce, W1 * mse, W2 * fa, (sssr, sisr, None, None)
        sssr, sisr, sssr_t, sisr_t = self.model(img)
        ce = self.ce(sssr, target if self.ce_takes_uint8 else target.long())
        mse = W1 * self.mse(sisr, org)
        fa = W2 * self.fa(sssr_t, sisr_t)
        return ce, mse, fa, (sssr, sisr, sssr_t, sisr_t)

    def __call__(self, img, org, target):
        ce, mse, fa, _ = self.losses(img, org, target)
        total = ce + mse + fa
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        self.opt.step()
        return ce.detach(), mse.detach(), fa.detach(), total.detach()
