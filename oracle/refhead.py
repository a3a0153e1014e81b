"""Oracle (TEST INFRASTRUCTURE): a module with the LAYOUT of the reference detection head, so that a ``state_dict`` saved
from the reference's own ``Head`` (training/yolopt/nets/nn.py:228-253) loads with ``strict=True`` on a box that has no
/root/reference.  Only the conv stacks are restated (nn.py:28-37 ``Conv`` = Conv2d(bias=False) + BatchNorm2d(eps=1e-3,
momentum=0.03) + SiLU; nn.py:243-253 the ``box`` / ``cls`` stacks; nn.py:215-220 the fixed ``dfl.conv`` weight); the eval
branch after them (nn.py:255-270) is what ``spp.head_eval_forward`` replaces and is NOT restated here."""
from __future__ import annotations

import torch


class _Conv(torch.nn.Module):
    def __init__(self, cin, cout, k=1, p=0, g=1):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, k, 1, p, groups=g, bias=False)
        self.norm = torch.nn.BatchNorm2d(cout, eps=0.001, momentum=0.03)
        self.relu = torch.nn.SiLU()

    def forward(self, x):
        return self.relu(self.norm(self.conv(x)))


class _Dfl(torch.nn.Module):
    def __init__(self, ch=16):
        super().__init__()
        self.conv = torch.nn.Conv2d(ch, 1, 1, bias=False).requires_grad_(False)


class RefShapedHead(torch.nn.Module):
    """``box`` / ``cls`` ModuleLists + ``stride`` exactly as the reference names them."""

    def __init__(self, nc: int, filters, stride=(8.0, 16.0, 32.0)):
        super().__init__()
        self.nc, self.ch = nc, 16
        box = max(64, filters[0] // 4)
        cls = max(80, filters[0], nc)
        self.dfl = _Dfl(self.ch)
        self.box = torch.nn.ModuleList(torch.nn.Sequential(_Conv(x, box, 3, 1), _Conv(box, box, 3, 1),
                                                           torch.nn.Conv2d(box, 4 * self.ch, 1)) for x in filters)
        self.cls = torch.nn.ModuleList(torch.nn.Sequential(_Conv(x, x, 3, 1, x), _Conv(x, cls), _Conv(cls, cls, 3, 1, cls),
                                                           _Conv(cls, cls), torch.nn.Conv2d(cls, nc, 1)) for x in filters)
        self.stride = torch.tensor(stride)
