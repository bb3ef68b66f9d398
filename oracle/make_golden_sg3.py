"""Generate tests/golden/filtered_lrelu.pt by running the REFERENCE's own ``_filtered_lrelu_ref`` (forward and, through
autograd, the gradients w.r.t. x and b) in the build container.  Usage: python -m oracle.make_golden_sg3"""
import os
import sys
import types

import torch

REF = '/root/reference'
CASES = [   # H, W, fu taps, fd taps, up, down, padding, clamp, bias, flip_filter
    dict(H=16, W=16, fu=12, fd=12, up=2, down=2, padding=10, clamp=None, bias=True, flip=False),       # SG3-T layer shape
    dict(H=13, W=9, fu=12, fd=12, up=2, down=2, padding=[11, 9, 10, 12], clamp=0.8, bias=True, flip=False),
    dict(H=12, W=12, fu=24, fd=12, up=4, down=2, padding=19, clamp=256.0, bias=True, flip=True),        # up 4 (SG3 critical layers)
    dict(H=10, W=14, fu=0, fd=0, up=1, down=1, padding=0, clamp=None, bias=True, flip=False),          # 1x1 / ToRGB case
    dict(H=20, W=20, fu=12, fd=0, up=2, down=1, padding=5, clamp=None, bias=False, flip=False),
    dict(H=20, W=20, fu=0, fd=12, up=1, down=2, padding=6, clamp=0.5, bias=False, flip=False),
    dict(H=18, W=18, fu=6, fd=6, up=2, down=2, padding=[-1, 3, 2, -2], clamp=0.5, bias=True, flip=False),   # outer(f, f) filters, crop
]


def main(out='tests/golden/filtered_lrelu.pt'):
    sys.path[:0] = [REF, os.path.join(REF, 'models/stylegan3')]
    for name in ['openpyxl', 'matplotlib', 'matplotlib.pyplot']:
        sys.modules.setdefault(name, types.ModuleType(name))
    from torch_utils.ops import filtered_lrelu as ref_fl
    from torch_utils.ops import upfirdn2d as ref_up
    cases = []
    for i, c in enumerate(CASES):
        g = torch.Generator().manual_seed(100 + i)
        x = torch.randn([2, 3, c['H'], c['W']], generator=g, requires_grad=True)
        b = torch.randn([3], generator=g, requires_grad=True) if c['bias'] else None
        fu = ref_up.setup_filter((torch.rand(c['fu'], generator=g) + 0.1).tolist()) if c['fu'] else None
        fd = ref_up.setup_filter((torch.rand(c['fd'], generator=g) + 0.1).tolist()) if c['fd'] else None
        y = ref_fl._filtered_lrelu_ref(x, fu=fu, fd=fd, b=b, up=c['up'], down=c['down'], padding=c['padding'], gain=2 ** 0.5, slope=0.2,
                                       clamp=c['clamp'], flip_filter=c['flip'])
        gy = torch.randn(y.shape, generator=g)
        (y * gy).sum().backward()
        cases.append(dict(cfg=c, x=x.detach(), b=None if b is None else b.detach(), fu=fu, fd=fd, y=y.detach(), gy=gy,
                          gx=x.grad.clone(), gb=None if b is None else b.grad.clone()))
        print(i, c, tuple(y.shape))
    torch.save(cases, out)


if __name__ == '__main__':
    main()
