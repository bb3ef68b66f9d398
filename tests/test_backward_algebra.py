"""CPU check of the algebra the CUDA path implements, written with plain
torch ops and NO autograd, against autograd through the oracle generator:

  * the x2 up-sampling conv (conv_transpose2d stride 2 + 4x4 FIR, reference
    conv2d_resample.py:112-129) folded into per-phase 3x3 weights,
  * the manual backward-to-w: data gradient only (no weight-gradient conv), style gradients
    from the two fused reductions red_s / red_d, toRGB + skip-pyramid backward,
  * the bank-moment form of the pixel / latent criteria gradients.

This mirrors latentaugment_b200/csrc/{kernels.cu,tapgemm.cu} formula for formula.
"""
import math

import torch
import torch.nn.functional as F

from oracle import latent_aug as ola
from oracle import ops, synthetic

SQRT2 = math.sqrt(2.0)


def composite_weights(w, fir):
    """[O,I,3,3] -> [4 phases][3][3][O,I]; kernels.cu:prep_conv_weights_kernel."""
    O, I = w.shape[:2]
    out = torch.zeros(4, 3, 3, O, I, dtype=w.dtype)
    for py in range(2):
        for px in range(2):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    for ay in range(3):
                        jy = ay + 1 - py + 2 * dy
                        if not 0 <= jy <= 3:
                            continue
                        for ax in range(3):
                            jx = ax + 1 - px + 2 * dx
                            if not 0 <= jx <= 3:
                                continue
                            out[py * 2 + px, dy + 1, dx + 1] += fir[3 - jy, 3 - jx] * 4.0 * w[:, :, ay, ax]
    return out


def shift(x, dy, dx):
    """y[p] = x[p + (dy,dx)] with zero padding; x [B,C,H,W]."""
    H, W = x.shape[2:]
    xp = F.pad(x, [1, 1, 1, 1])
    return xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]


def conv_fwd(xs, w, up, fir):
    """tap-GEMM forward: normal 3x3 (taps a-1) or 4 phase problems with composite weights."""
    B, I, H, W = xs.shape
    O = w.shape[0]
    if up == 1:
        y = torch.zeros(B, O, H, W, dtype=xs.dtype)
        for ay in range(3):
            for ax in range(3):
                y += torch.einsum('bihw,oi->bohw', shift(xs, ay - 1, ax - 1), w[:, :, ay, ax])
        return y
    wc = composite_weights(w, fir)
    y = torch.zeros(B, O, 2 * H, 2 * W, dtype=xs.dtype)
    for ph in range(4):
        acc = torch.zeros(B, O, H, W, dtype=xs.dtype)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                acc += torch.einsum('bihw,oi->bohw', shift(xs, dy, dx), wc[ph, dy + 1, dx + 1])
        y[:, :, ph // 2::2, ph % 2::2] = acc
    return y


def conv_dgrad(gy, w, up, fir):
    """tap-GEMM data gradient."""
    if up == 1:
        g = torch.zeros(gy.shape[0], w.shape[1], *gy.shape[2:], dtype=gy.dtype)
        for ay in range(3):
            for ax in range(3):
                g += torch.einsum('bohw,oi->bihw', shift(gy, 1 - ay, 1 - ax), w[:, :, ay, ax])
        return g
    wc = composite_weights(w, fir)
    H, W = gy.shape[2] // 2, gy.shape[3] // 2
    g = torch.zeros(gy.shape[0], w.shape[1], H, W, dtype=gy.dtype)
    for ph in range(4):
        plane = gy[:, :, ph // 2::2, ph % 2::2]
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                g += torch.einsum('bohw,oi->bihw', shift(plane, -dy, -dx), wc[ph, dy + 1, dx + 1])
    return g


def up2(img):
    """rgb_combine_kernel's skip up-sampling: out[2m] = (x[m-1]+3x[m])/4, out[2m+1] = (3x[m]+x[m+1])/4."""
    def up1d(x, dim):
        n = x.shape[dim]
        xm = torch.cat([torch.zeros_like(x.narrow(dim, 0, 1)), x.narrow(dim, 0, n - 1)], dim)
        xp = torch.cat([x.narrow(dim, 1, n - 1), torch.zeros_like(x.narrow(dim, 0, 1))], dim)
        even, odd = 0.25 * xm + 0.75 * x, 0.75 * x + 0.25 * xp
        return torch.stack([even, odd], dim + 1).flatten(dim, dim + 1)
    return up1d(up1d(img, 2), 3)


def up2_T(g):
    """rgb_backward_kernel: low-res m gets out[2m-1]/4 + 3 out[2m]/4 + 3 out[2m+1]/4 + out[2m+2]/4."""
    def down1d(x, dim):
        n = x.shape[dim]
        xp = F.pad(x.movedim(dim, -1), [1, 1]).movedim(-1, dim)
        idx = torch.arange(0, n, 2)
        return (0.25 * xp.index_select(dim, idx) + 0.75 * xp.index_select(dim, idx + 1) +
                0.75 * xp.index_select(dim, idx + 2) + 0.25 * xp.index_select(dim, idx + 3))
    return down1d(down1d(g, 2), 3)


def manual_step_gradient(G, w, W_bank, X_bank, w_latent=1.0, w_pix=1.0, dtype=torch.float64):
    """Forward + manual backward exactly as the CUDA engine does it.  Returns (img, dL/dw, l_lat, l_pix)."""
    S = G.synthesis
    B = w.shape[0]
    w = w.to(dtype)
    convs, rgbs = [], []
    for r in S.block_resolutions:
        blk = getattr(S, f'b{r}')
        if r > 4:
            convs.append((blk.conv0, 2, r))
        convs.append((blk.conv1, 1, r))
        rgbs.append((blk.torgb, r, len(convs) - 1))
    fir = S.b4.resample_filter.to(dtype)
    P = lambda t: t.detach().to(dtype)
    aff = lambda a: (P(a.weight) * a.w_gain, P(a.bias) * a.b_gain)
    # ---- styles / demod
    s, d, W2 = [], [], []
    for L, up, r in convs:
        A, b = aff(L.affine)
        s.append(w @ A.t() + b)
        W2.append(P(L.weight).square().sum([2, 3]))
        d.append((s[-1].square() @ W2[-1].t() + 1e-8).rsqrt())
    s_rgb = []
    for T, r, li in rgbs:
        A, b = aff(T.affine)
        s_rgb.append((w @ A.t() + b) * T.w_gain)
    # ---- forward
    x_prev = P(S.b4.const).unsqueeze(0).expand(B, -1, -1, -1)
    xs_in, x_out, noise = [], [], []
    for l, (L, up, r) in enumerate(convs):
        xs = x_prev * s[l][:, :, None, None]
        y = conv_fwd(xs, P(L.weight), up, fir)
        nz = P(L.noise_const) * P(L.noise_strength)
        z = y * d[l][:, :, None, None] + nz + P(L.bias)[None, :, None, None]
        out = torch.where(z > 0, z, 0.2 * z) * SQRT2
        out = out.clamp(-L.conv_clamp, L.conv_clamp)
        noise.append(nz)
        x_out.append(out)
        x_prev = out
    img, pre = None, []
    for k, (T, r, li) in enumerate(rgbs):
        t = torch.einsum('bihw,ci,bi->bchw', x_out[li], P(T.weight)[:, :, 0, 0], s_rgb[k]) + P(T.bias)[None, :, None, None]
        pre.append(t)
        y = t.clamp(-T.conv_clamp, T.conv_clamp)
        img = y if img is None else up2(img) + y
    # ---- criteria (bank moments)
    res, C = G.img_resolution, G.img_channels
    off, size = ola.center_crop_bounds(res)
    ybar = X_bank.to(dtype).mean(0)
    m2 = X_bank.to(dtype)[:, :, off:off + size, off:off + size].square().sum([2, 3]).mean(0)      # [C]
    xc = img[:, :, off:off + size, off:off + size]
    yc = ybar[:, off:off + size, off:off + size]
    l_pix = w_pix * ((xc.square() - 2 * xc * yc).sum() / (B * size * size) + m2.sum() / (size * size)) / C
    g_img = torch.zeros_like(img)
    g_img[:, :, off:off + size, off:off + size] = -(w_pix / C) * 2.0 / (B * size * size) * (xc - yc)
    num_ws, w_dim = G.num_ws, G.w_dim
    Wb = W_bank.to(dtype)
    w_sum, lat_m2 = Wb.mean(0).sum(0), Wb.square().sum([1, 2]).mean()
    l_lat = w_latent * ((num_ws * w.square().sum() - 2 * (w @ w_sum).sum()) / B + lat_m2) / (num_ws * w_dim)
    g_w = -w_latent * 2.0 / (B * num_ws * w_dim) * (num_ws * w - w_sum)
    # ---- toRGB / skip pyramid backward (top-down)
    g_rgb = [None] * len(rgbs)
    g = g_img
    for k in reversed(range(len(rgbs))):
        T = rgbs[k][0]
        g_rgb[k] = g * (pre[k].abs() <= T.conv_clamp)
        if k > 0:
            g = up2_T(g)
    # ---- conv chain backward
    rgb_of_layer = {li: k for k, (T, r, li) in enumerate(rgbs)}
    g_s = [None] * len(convs)
    red_d = [None] * len(convs)
    g_s_rgb = [None] * len(rgbs)

    def act_backward(l, g_x):
        L = convs[l][0]
        out = x_out[l]
        pos = out > 0
        gz = g_x * SQRT2 * torch.where(pos, 1.0, 0.2) * (out.abs() < L.conv_clamp)
        z = torch.where(pos, out / SQRT2, out / (SQRT2 * 0.2))
        red_d[l] = (gz * (z - noise[l] - P(L.bias)[None, :, None, None])).sum([2, 3])
        return gz * d[l][:, :, None, None]

    top = len(convs) - 1
    k = rgb_of_layer[top]
    T = rgbs[k][0]
    rgbw = P(T.weight)[:, :, 0, 0][None] * s_rgb[k][:, None, :]                  # [B,C,I]
    g_s_rgb[k] = torch.einsum('bihw,bchw,ci->bi', x_out[top], g_rgb[k], P(T.weight)[:, :, 0, 0])
    gy = act_backward(top, torch.einsum('bchw,bci->bihw', g_rgb[k], rgbw))     # seed
    for l in range(top, -1, -1):
        L, up, r = convs[l]
        g_xs = conv_dgrad(gy, P(L.weight), up, fir)
        x_in = x_out[l - 1] if l > 0 else P(S.b4.const).unsqueeze(0).expand(B, -1, -1, -1)
        red_s = (g_xs * x_in).sum([2, 3])
        g_s[l] = red_s - s[l] * ((red_d[l] * d[l].square()) @ W2[l])
        if l == 0:
            break
        g_x = g_xs * s[l][:, :, None, None]
        if (l - 1) in rgb_of_layer:
            k = rgb_of_layer[l - 1]
            T = rgbs[k][0]
            rgbw = P(T.weight)[:, :, 0, 0][None] * s_rgb[k][:, None, :]
            g_x = g_x + torch.einsum('bchw,bci->bihw', g_rgb[k], rgbw)
            g_s_rgb[k] = torch.einsum('bihw,bchw,ci->bi', x_out[l - 1], g_rgb[k], P(T.weight)[:, :, 0, 0])
        gy = act_backward(l - 1, g_x)
    # ---- styles -> w
    for l, (L, up, r) in enumerate(convs):
        A, _ = aff(L.affine)
        g_w = g_w + g_s[l] @ A
    for k, (T, r, li) in enumerate(rgbs):
        A, _ = aff(T.affine)
        g_w = g_w + (g_s_rgb[k] * T.w_gain) @ A
    return img, g_w, float(l_lat), float(l_pix)


def test_composite_upconv_equals_reference_composition():
    gen = torch.Generator().manual_seed(0)
    x = torch.randn([2, 5, 6, 6], generator=gen, dtype=torch.float64)
    w = torch.randn([4, 5, 3, 3], generator=gen, dtype=torch.float64)
    f = ops.setup_filter([1, 3, 3, 1]).double()
    ref = ops.conv2d_resample(x, w, f=f, up=2, padding=1, flip_weight=False)
    out = conv_fwd(x, w, 2, f)
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)
    ref1 = ops.conv2d_resample(x, w, f=None, up=1, padding=1, flip_weight=True)
    torch.testing.assert_close(conv_fwd(x, w, 1, f), ref1, rtol=1e-12, atol=1e-12)
    # data gradients = adjoint
    gy = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(ops.conv2d_resample(xr, w, f=f, up=2, padding=1, flip_weight=False), xr, gy)
    torch.testing.assert_close(conv_dgrad(gy, w, 2, f), gx, rtol=1e-12, atol=1e-12)
    gy1 = torch.randn(ref1.shape, generator=gen, dtype=torch.float64)
    (gx1,) = torch.autograd.grad(ops.conv2d_resample(xr, w, f=None, up=1, padding=1, flip_weight=True), xr, gy1)
    torch.testing.assert_close(conv_dgrad(gy1, w, 1, f), gx1, rtol=1e-12, atol=1e-12)


def test_skip_upsample_equals_reference_upsample2d():
    x = torch.randn([2, 3, 8, 8], generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    f = ops.setup_filter([1, 3, 3, 1]).double()
    torch.testing.assert_close(up2(x), ops.upsample2d(x, f), rtol=1e-12, atol=1e-12)
    xr = x.clone().requires_grad_(True)
    g = torch.randn([2, 3, 16, 16], generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    (gx,) = torch.autograd.grad(ops.upsample2d(xr, f), xr, g)
    torch.testing.assert_close(up2_T(g), gx, rtol=1e-12, atol=1e-12)


def test_manual_backward_to_w_equals_autograd():
    wl = synthetic.make_workload('tiny', noise_strength=0.1)
    G = wl['G'].double()
    w0 = wl['w0'][:, 0].double()
    img, g_w, l_lat, l_pix = manual_step_gradient(G, w0, wl['W'], wl['X'])
    w = w0.clone().requires_grad_(True)
    ws = w.unsqueeze(1).repeat(1, G.num_ws, 1)
    S, xx, x, idx = G.synthesis, None, None, 0          # Synthesis.forward without its float32 cast
    for r in S.block_resolutions:
        blk = getattr(S, f'b{r}')
        xx, x = blk(xx, x, ws.narrow(1, idx, blk.num_conv + blk.num_torgb), noise_mode='const', fused=True)
        idx += blk.num_conv
    x = x.double()                     # the block casts the toRGB output to float32 (as upstream does)
    ll = ola.calc_loss_latent(ws, wl['W'].double(), 1.0)
    lp = ola.calc_loss_pix(ola.center_crop(x, G.img_resolution), ola.center_crop(wl['X'].double(), G.img_resolution), 1.0,
                           G.img_channels)
    (g_ref,) = torch.autograd.grad(-ll - lp, w)
    torch.testing.assert_close(img, x.detach(), rtol=1e-5, atol=1e-5)
    assert abs(l_lat - float(ll)) < 1e-9 * abs(float(ll)) and abs(l_pix - float(lp)) < 1e-6 * abs(float(lp))
    rel = float((g_w - g_ref).norm() / g_ref.norm())
    assert rel < 1e-5, rel


def test_bank_moment_forms_equal_the_pairwise_losses():
    """DESIGN.md §3.3 / SURVEY.md App. B: the loop keeps only the banks' first and second moments.  Values and gradients of
    the latent, pixel and perceptual criteria computed from moments (the formulas the kernels implement, in fp64) against
    autograd through the oracle's pairwise forms."""
    from oracle import latent_aug as ola
    from oracle import lpips as olp
    g = torch.Generator().manual_seed(3)
    # ---- latent: w_latent * mean_ij |x_i - y_j|^2 / K
    n, m, nw, K = 5, 37, 6, 16
    x = torch.randn([n, nw, K], generator=g, dtype=torch.float64, requires_grad=True)
    Y = torch.randn([m, nw, K], generator=g, dtype=torch.float64)
    loss = ola.calc_loss_latent(x, Y, 0.7)
    loss.backward()
    xf, Yf = x.detach().reshape(n, -1), Y.reshape(m, -1)
    ybar, y2 = Yf.mean(0), (Yf * Yf).sum(1).mean()
    val = 0.7 * ((xf * xf).sum(1).mean() - 2.0 * (xf.mean(0) * ybar).sum() + y2) / (nw * K)
    grad = 0.7 * 2.0 / (n * nw * K) * (xf - ybar)
    assert abs(float(val - loss.detach())) < 1e-12 * abs(float(loss.detach()))
    assert torch.allclose(grad.reshape(n, nw, K), x.grad, rtol=1e-12, atol=1e-15)
    # ---- pixel: per modality over the centre crop, averaged over modalities
    C, res = 2, 32
    img = torch.randn([n, C, res, res], generator=g, dtype=torch.float64, requires_grad=True)
    X = torch.randn([m, C, res, res], generator=g, dtype=torch.float64)
    loss = ola.calc_loss_pix(ola.center_crop(img, res), ola.center_crop(X, res), 0.3, C)
    loss.backward()
    off, size = ola.center_crop_bounds(res)
    val, grad = 0.0, torch.zeros_like(img)
    for c in range(C):
        xc = img.detach()[:, c, off:off + size, off:off + size].reshape(n, -1)
        bc = X[:, c, off:off + size, off:off + size].reshape(m, -1)
        val = val + 0.3 * ((xc * xc).sum(1).mean() - 2.0 * (xc.mean(0) * bc.mean(0)).sum() + (bc * bc).sum(1).mean()) / (size * size) / C
        grad[:, c, off:off + size, off:off + size] = (0.3 * 2.0 / (n * size * size * C) * (xc - bc.mean(0))).reshape(n, size, size)
    assert abs(float(val - loss.detach())) < 1e-12 * abs(float(loss.detach()))
    assert torch.allclose(grad, img.grad, rtol=1e-12, atol=1e-15)
    # ---- perceptual: pair distance sum_taps mean_hw sum_c w_c (n^_x - n^_y)^2 through the bank moments b = mean_j n^_j and
    # M2 = mean_j sum_c w_c n^_j^2 / hw; gradient w.r.t. the UN-normalised activation through the normalisation pull-back
    taps = olp.TAPS_INTREE
    st = {k: v.double() for k, v in olp.random_vgg_state(7, taps).items()}
    lw = olp.lin_weights(st, taps)
    acts = [torch.rand([n, c, s, s], generator=g, dtype=torch.float64, requires_grad=True) for c, s in ((256, 4), (512, 2), (512, 1))]
    bank = [olp.normalize_activation(torch.rand([m, a.shape[1], a.shape[2], a.shape[3]], generator=g, dtype=torch.float64)) for a in acts]
    for script in (True, False):
        for a in acts:
            a.grad = None
        D = olp.pair_distance([olp.normalize_activation(a) for a in acts], bank, lw)          # [m, n]
        loss = 2.5 * (D.sum() / (n * m) if script else D.sum() / m)
        loss.backward()
        norm = 1.0 / n if script else 1.0                                # per-pair weight after the bank mean: 1/(n m) or 1/m
        val = 0.0
        for a, b, w in zip(acts, bank, lw):
            x = a.detach()
            hw = x.shape[2] * x.shape[3]
            wv = w.reshape(1, -1, 1, 1)
            r = torch.sqrt((x * x).sum(1, keepdim=True))
            nh = x / (r + 1e-10)
            bbar = b.mean(0, keepdim=True)
            m2 = (wv * b * b).sum((1, 2, 3)).mean() / hw
            val = val + 2.5 * norm * ((wv * (nh * nh - 2.0 * nh * bbar)).sum() / hw + n * m2)
            seed = 2.5 * norm * (2.0 / hw) * wv * (nh - bbar)                                  # d loss / d n^
            gx = seed / (r + 1e-10) - x * (seed * x).sum(1, keepdim=True) / (r * (r + 1e-10) ** 2)
            assert torch.allclose(gx, a.grad, rtol=1e-9, atol=1e-14), script
        assert abs(float(val - loss.detach())) < 1e-11 * abs(float(loss.detach())), script
