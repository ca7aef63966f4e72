"""Dense, differentiable PyTorch (float64) model of the rasterizer -- an INDEPENDENT cross-check
of oracle/gms_oracle.c (forward compositing and every gradient, via autograd).

TEST INFRASTRUCTURE ONLY.  O(pixels x Gaussians) memory: use for P <~ 500 and images <~ 64x64.

It follows SURVEY.md Appendix A; the pieces that the reference ships in Python are restated from
  utils/sh_utils.py:57-112 (SH basis), utils/general_utils.py:158-190 (R from quaternion, R*S),
  utils/graphics_utils.py:22-29 (homogeneous divide with +1e-7).
Integer decisions (radius -> tile rectangle, near cull) are taken from the C oracle's forward state
so that the two models composite exactly the same per-tile lists; all continuous maths is redone here.
Gradient conventions replicated from the stock rasterizer (Appendix A.3/A.4):
  * alpha clamp at 0.99 passes gradient straight through,
  * the guard-band clamp of t.x/t.z, t.y/t.z zeroes only the x / y term,
  * means2D receives the NDC-scaled screen-space gradient (0.5*W, 0.5*H).
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def sh_to_rgb(deg, sh, d):
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
               + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
               + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def quat_to_R(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.reshape(-1, 3, 3)


class _StraightThroughMin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cap):
        return torch.clamp(x, max=cap)

    @staticmethod
    def backward(ctx, g):
        return g, None


def render(settings, rects, means3D, means2D_sink, opacities, shs=None, colors_precomp=None,
           scales=None, rotations=None, cov3D_precomp=None, depths32=None):
    """settings: oracle.raster.Settings; rects: [P,4] int (xmin,ymin,xmax,ymax tiles; empty => invisible)
    taken from the C oracle.  All tensor inputs float64 (requires_grad as desired).
    `means2D_sink` [P,3] is the gradient sink of the API: it is added (in NDC-scaled units) to the pixel
    centre so that its autograd gradient equals the stock dL_dmean2D convention.
    `depths32` (optional) = the C oracle's fp32 depths, used only for the sort order.
    Returns color [3,H,W], invdepth [1,H,W], final_T [H,W], n_contrib [H,W]."""
    dt = torch.float64
    W, H = settings.image_width, settings.image_height
    view = torch.as_tensor(settings.viewmatrix, dtype=dt).reshape(4, 4)   # transposed storage
    proj = torch.as_tensor(settings.projmatrix, dtype=dt).reshape(4, 4)
    campos = torch.as_tensor(settings.campos, dtype=dt)
    bg = torch.as_tensor(settings.bg, dtype=dt)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dt)
    hom = torch.cat([means3D, ones], 1)
    p_view = (hom @ view)[:, :3]
    p_hom = hom @ proj
    p_w = 1.0 / (p_hom[:, 3:4] + 1e-7)
    p_proj = p_hom[:, :3] * p_w
    # cov3D
    if cov3D_precomp is None:
        R = quat_to_R(rotations)
        Mx = R * (settings.scale_modifier * scales)[:, None, :]
        Sigma = Mx @ Mx.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], 1).reshape(-1, 3, 3)
    fx = W / (2.0 * settings.tanfovx)
    fy = H / (2.0 * settings.tanfovy)
    tz = p_view[:, 2]
    limx, limy = 1.3 * settings.tanfovx, 1.3 * settings.tanfovy
    txtz, tytz = p_view[:, 0] / tz, p_view[:, 1] / tz
    # stock semantics: clamp -> the x/y gradient term is zeroed; tz is still differentiated as if unclamped
    inx = ((txtz >= -limx) & (txtz <= limx)).to(dt)
    iny = ((tytz >= -limy) & (tytz <= limy)).to(dt)
    tx_c = torch.clamp(txtz, -limx, limx).detach() * tz.detach()
    ty_c = torch.clamp(tytz, -limy, limy).detach() * tz.detach()
    tx = tx_c + inx * (p_view[:, 0] - p_view[:, 0].detach())
    ty = ty_c + iny * (p_view[:, 1] - p_view[:, 1].detach())
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / (tz * tz), zero, fy / tz, -fy * ty / (tz * tz)], 1).reshape(-1, 2, 3)
    Wm = view[:3, :3].t()            # true rotation: W[i][j] = view_flat[4*j+i]
    Mj = J @ Wm
    cov2 = Mj @ Sigma @ Mj.transpose(1, 2)
    a0, b, c0 = cov2[:, 0, 0], cov2[:, 0, 1], cov2[:, 1, 1]
    det_cov = a0 * c0 - b * b
    a, c = a0 + 0.3, c0 + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], 1)
    if settings.antialiasing:
        h = torch.sqrt(torch.clamp(det_cov / det, min=0.000025))
    else:
        h = torch.ones_like(det)
    op = opacities.reshape(-1) * h
    pix = torch.stack([((p_proj[:, 0] + 1.0) * W - 1.0) * 0.5, ((p_proj[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    pix = pix + means2D_sink[:, :2] * torch.tensor([0.5 * W, 0.5 * H], dtype=dt)
    # colour
    if colors_precomp is None:
        d = means3D - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb_raw = sh_to_rgb(settings.sh_degree, shs, d) + 0.5
        rgb = torch.clamp(rgb_raw, min=0.0)
    else:
        rgb = colors_precomp
    invd = 1.0 / tz

    rects = torch.as_tensor(rects, dtype=torch.int64)
    vis = (rects[:, 2] > rects[:, 0]) & (rects[:, 3] > rects[:, 1])
    # global depth order (fp32 depth bits order == fp32 value order for positive floats); ties by index
    depth32 = p_view[:, 2].detach().to(torch.float32) if depths32 is None else torch.as_tensor(depths32)
    order = torch.argsort(depth32, stable=True)
    order = order[vis[order]]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pxf = xs.reshape(-1).to(dt); pyf = ys.reshape(-1).to(dt)
    tix = (xs.reshape(-1) // 16); tiy = (ys.reshape(-1) // 16)
    G = order.numel()
    color = torch.zeros(3, H * W, dtype=dt)
    if G == 0:
        out = color + bg[:, None]
        return out.reshape(3, H, W), torch.zeros(1, H, W, dtype=dt), torch.ones(H, W, dtype=dt), torch.zeros(H, W, dtype=torch.int64)
    r = rects[order]
    in_list = (tix[:, None] >= r[None, :, 0]) & (tix[:, None] < r[None, :, 2]) & \
              (tiy[:, None] >= r[None, :, 1]) & (tiy[:, None] < r[None, :, 3])       # [Npix, G]
    dx = pix[order, 0][None] - pxf[:, None]
    dy = pix[order, 1][None] - pyf[:, None]
    con = conic[order]
    power = -0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) - con[None, :, 1] * dx * dy
    Gv = torch.exp(power)
    alpha = _StraightThroughMin.apply(op[order][None] * Gv, 0.99)
    contrib = in_list & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
    a_eff = torch.where(contrib, alpha, torch.zeros_like(alpha))
    one_m = 1.0 - a_eff
    Tincl = torch.cumprod(one_m, dim=1)                      # T after each splat
    Tbefore = torch.cat([torch.ones(H * W, 1, dtype=dt), Tincl[:, :-1]], 1)
    stop = contrib & (Tincl.detach() < 0.0001)               # test_T < 1e-4 -> done, splat not blended
    stopped = torch.cumsum(stop.to(torch.int64), dim=1) > 0
    live = contrib & ~stopped
    wgt = torch.where(live, a_eff * Tbefore, torch.zeros_like(a_eff))
    color = (wgt @ rgb[order]).t()                            # [3, Npix]
    # final T = product over live splats
    logs = torch.where(live, one_m, torch.ones_like(one_m))
    final_T = torch.prod(logs, dim=1)
    color = color + final_T[None] * bg[:, None]
    invdepth = (wgt @ invd[order][:, None]).t()
    # n_contrib: 1-based position of the last live splat within the pixel's tile list
    pos = torch.cumsum(in_list.to(torch.int64), dim=1)
    n_contrib = torch.where(live, pos, torch.zeros_like(pos)).max(dim=1).values
    return color.reshape(3, H, W), invdepth.reshape(1, H, W), final_T.reshape(H, W), n_contrib.reshape(H, W)
