"""Shared pieces of the three networks: the residual stack with its running skip sum and the
LeakyReLU/1x1/LeakyReLU/1x1 output head."""
import torch.nn as nn

from .. import functional as WF


def make_output_head(out_dim, n_out):
    """Parameter layout of the reference heads (`output_stack` wavenet.py:67-71, `output_block`
    raw_ctcnet.py:84-88 / classifier.py:70-74): indices 1 and 3 hold the 1x1 convs."""
    return nn.Sequential(nn.LeakyReLU(0.01), nn.Conv1d(out_dim, out_dim, kernel_size=1, padding=0, dilation=1),
                         nn.LeakyReLU(0.01), nn.Conv1d(out_dim, n_out, kernel_size=1, padding=0, dilation=1))


def run_head(head, skips):
    return WF.output_stack(skips, head[1].weight, head[1].bias, head[3].weight, head[3].bias)


def run_stack(out, skips, blocks, bottlenecks):
    """for l: out, skip = block_l(out); skips += bottleneck_l(skip)  (wavenet.py:98-100)."""
    for block, bn in zip(blocks, bottlenecks):
        out, skip = block(out)
        skips = WF.skip_accumulate(skips, skip, bn.weight, bn.bias)
    return out, skips


def kaiming_weights_(params, bias_fill):
    """kaiming-uniform on every >=2-D parameter, `bias_fill(p)` on every 1-D one (the init loop the three
    reference constructors share)."""
    for p in params:
        if p.dim() > 1:
            nn.init.kaiming_uniform_(p)
        if p.dim() == 1:
            bias_fill(p)
