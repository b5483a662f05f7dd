"""Oracle for R2, the differentiable sphere depth renderer.  TEST INFRASTRUCTURE (see oracle/__init__).

Restates, in numpy fp32 with the reference's exact operation order:
  * BallRender.forward                      /root/reference/mesh/render.py:26-53
  * min over the J spheres of one image     /root/reference/mesh/render.py:89,
                                            /root/reference/mesh/multiview_utility.py:76
and the analytic backward that torch autograd derives from them.
"""
import numpy as np

F32 = np.float32
BACKGROUND = F32(100.0)
S_MIN = F32(1e-2)


def pixel_grid(width, height):
    """xg(u) = (u - W/2) * 300 / W ; yg(v) alike.  render.py:31-32 (no half-pixel offset)."""
    u = np.arange(width, dtype=F32)
    v = np.arange(height, dtype=F32)
    xg = (u - F32(width / 2)) * F32(300.0) / F32(width)
    yg = (v - F32(height / 2)) * F32(300.0) / F32(height)
    return xg.astype(F32), yg.astype(F32)


def ball_render(centres, radii, width, height):
    """One depth image per sphere.  centres [N,>=3], radii [N] -> [N,H,W] fp32.  render.py:26-53."""
    centres = np.asarray(centres, dtype=F32)
    radii = np.asarray(radii, dtype=F32)
    xg, yg = pixel_grid(width, height)
    cx = centres[:, 0].reshape(-1, 1, 1)
    cy = centres[:, 1].reshape(-1, 1, 1)
    cz = centres[:, 2].reshape(-1, 1, 1)
    r = radii.reshape(-1, 1, 1)
    xs = (xg.reshape(1, 1, -1) - cx) ** 2           # render.py:37
    ys = (yg.reshape(1, -1, 1) - cy) ** 2           # render.py:38
    s = (r * r - xs) - ys                           # render.py:41 (left-to-right)
    s = np.maximum(s, S_MIN)                        # clamp(min=1e-2)
    fg = s != S_MIN                                 # render.py:42
    depth = np.where(fg, cz - np.sqrt(s), BACKGROUND).astype(F32)   # render.py:47,52
    return depth


def sphere_render(centres, radii, width, height):
    """centres [N,J,>=3], radii [J] or [N,J] -> depth [N,H,W] fp32, idx [N,H,W] uint8 (255 = background).

    depth = min_k d_k, idx = argmin_k d_k (first minimum), render.py:89."""
    centres = np.asarray(centres, dtype=F32)
    n, j = centres.shape[:2]
    radii = np.broadcast_to(np.asarray(radii, dtype=F32), (n, j))
    parts = ball_render(centres.reshape(n * j, -1), radii.reshape(-1), width, height)
    parts = parts.reshape(n, j, height, width)
    idx = parts.argmin(axis=1)
    depth = np.take_along_axis(parts, idx[:, None], axis=1)[:, 0]
    idx = np.where(depth < BACKGROUND, idx, 255).astype(np.uint8)
    return depth, idx


def sphere_render_tie_mask(centres, radii, width, height, tol=0.0):
    """True where two foreground spheres give the same (within tol) minimum depth (arg-min ambiguous)."""
    centres = np.asarray(centres, dtype=F32)
    n, j = centres.shape[:2]
    radii = np.broadcast_to(np.asarray(radii, dtype=F32), (n, j))
    parts = ball_render(centres.reshape(n * j, -1), radii.reshape(-1), width, height).reshape(n, j, height, width)
    mn = parts.min(axis=1, keepdims=True)
    return ((parts <= mn + F32(tol)).sum(axis=1) > 1) & (mn[:, 0] < BACKGROUND)


def sphere_render_backward(grad_depth, idx, centres, radii, width, height):
    """Analytic gradient: per-pixel terms from the fp32 forward values (as autograd saves them), summed in
    fp64.  Returns grad_centres [N,J,3], grad_radii [N,J].

    dd/dcx = -(xg-cx)/sqrt(s), dd/dcy alike, dd/dcz = 1, dd/dr = -r/sqrt(s); only the arg-min sphere of
    a foreground pixel receives gradient."""
    centres = np.asarray(centres, dtype=F32)
    n, j = centres.shape[:2]
    radii = np.broadcast_to(np.asarray(radii, dtype=F32), (n, j))
    xg, yg = pixel_grid(width, height)
    gc = np.zeros((n, j, 3))
    gr = np.zeros((n, j))
    g = np.asarray(grad_depth, dtype=np.float64)
    for i in range(n):
        vv, uu = np.nonzero(idx[i] != 255)
        k = idx[i][vv, uu].astype(np.int64)
        dx = xg[uu] - centres[i, k, 0]
        dy = yg[vv] - centres[i, k, 1]
        r = radii[i, k]
        sq = np.sqrt((r * r - dx * dx) - dy * dy).astype(np.float64)      # fp32 s, like the forward
        dx, dy, r = dx.astype(np.float64), dy.astype(np.float64), r.astype(np.float64)
        gg = g[i][vv, uu]
        np.add.at(gc[i, :, 0], k, gg * (-dx / sq))
        np.add.at(gc[i, :, 1], k, gg * (-dy / sq))
        np.add.at(gc[i, :, 2], k, gg)
        np.add.at(gr[i], k, gg * (-r / sq))
    return gc, gr
