"""oracle/dibr_oracle.c (the reference viewer's DIBR fragment shader, viewer.py:386-631, restated in C) against an INDEPENDENT numpy
restatement of the same shader written here from the GLSL text, plus known-answer properties.  No OpenGL context exists in the build
container and the reference never sets u_resolution, so this oracle is PARITY UNPINNED against the reference itself; what is
checked is that two separately written restatements agree and that the shader's documented behaviour holds."""
import math

import numpy as np
import pytest

from oracle import dibr

F = np.float32


def _tex(img, u, v):
    """texture(): GL_LINEAR, GL_REPEAT; img [h,w] or [h,w,c]; u,v arrays"""
    h, w = img.shape[:2]
    x = u * F(w) - F(0.5); y = v * F(h) - F(0.5)
    x0 = np.floor(x); y0 = np.floor(y)
    fx = (x - x0).astype(F); fy = (y - y0).astype(F)
    ix0 = np.mod(x0, w).astype(int); ix1 = np.mod(x0 + 1, w).astype(int)
    iy0 = np.mod(y0, h).astype(int); iy1 = np.mod(y0 + 1, h).astype(int)
    if img.ndim == 3:
        fx = fx[..., None]; fy = fy[..., None]
    top = img[iy0, ix0] * (F(1) - fx) + img[iy0, ix1] * fx
    bot = img[iy1, ix0] * (F(1) - fx) + img[iy1, ix1] * fx
    return (top * (F(1) - fy) + bot * fy).astype(F)


def _ss(e0, e1, x):
    t = np.clip((x - F(e0)) / (F(e1) - F(e0)), 0, 1).astype(F)
    return (t * t * (F(3) - F(2) * t)).astype(F)


def numpy_eye(color, depth, vw, vh, eye, strength, conv, res, search=12, tol=0.012, blur=2.5):
    """one eye view, vectorised over fragments; roll = 0, no feather, no corner radius"""
    j, i = np.meshgrid(np.arange(vw), np.arange(vh))
    uvx = ((j.astype(F) + F(0.5)) / F(vw)).astype(F)
    uvy = (((vh - 1 - i).astype(F) + F(0.5)) / F(vh)).astype(F)
    fx, fy = uvx, (F(1) - uvy).astype(F)
    psx, psy = F(1) / F(res[0]), F(1) / F(res[1])
    sg = F(np.sign(eye)); pdx, pdy = F(1) * sg, F(0) * sg
    sweep_sign = F(-1) if eye > 0 else F(1)
    dsx = pdx * psx * F(1.5)
    d0 = _tex(depth, fx, fy); dm = _tex(depth, fx - dsx, fy); dp = _tex(depth, fx + dsx, fy)
    dep = (d0 * F(0.7) + dm * F(0.15) + dp * F(0.15)).astype(F)
    dinv = -dep
    shaped = (dinv * (F(1) + F(0.35) * (F(1) - dep))).astype(F)
    shift = shaped + F(conv)
    fall = _ss(0.0, 0.05, fx) * _ss(1.0, F(1) - F(0.05), fx)
    px = (F(eye) * shift * F(strength) * fall).astype(F)
    sx, sy = (fx - px * F(1)).astype(F), (fy - px * F(0)).astype(F)
    s2 = pdx * psx * F(2)
    jump = np.abs(_tex(depth, fx - s2, fy) - _tex(depth, fx + s2, fy))
    conf = np.where((sx < 0) | (sx > 1) | (sy < 0) | (sy > 1), F(1), _ss(0.04, 0.10, jump)).astype(F)
    col = _tex(color, sx, sy)
    # push-pull inpaint for every fragment (selected by conf afterwards)
    w1, w2 = dibr.exp_tables()
    best = np.zeros(fx.shape + (3,), F); bw = np.zeros(fx.shape, F); done = np.zeros(fx.shape, bool)
    sw = pdx * psx * sweep_sign
    for k in range(1, search + 1):
        qx = (fx + sw * F(k)).astype(F)
        ok = ~done & ~((qx < 0) | (fy < 0) | (qx > 1) | (fy > 1))
        sdi = (F(1) - _tex(depth, qx, fy)).astype(F)
        acc = ok & (sdi > dinv + F(tol))
        wgt = (w1[k] * (F(1) + (sdi - dinv) * F(10))).astype(F)
        best = np.where(acc[..., None], best + _tex(color, qx, fy) * wgt[..., None], best).astype(F)
        bw = np.where(acc, bw + wgt, bw).astype(F)
        done |= acc & (bw > 5)
    ph2 = bw < 2
    for k in range(1, search + 1):
        qx = (fx - sw * F(k)).astype(F)
        ok = ph2 & ~((qx < 0) | (fy < 0) | (qx > 1) | (fy > 1))
        sdi = (F(1) - _tex(depth, qx, fy)).astype(F)
        acc = ok & (sdi > dinv + F(tol))
        best = np.where(acc[..., None], best + _tex(color, qx, fy) * w2[k], best).astype(F)
        bw = np.where(acc, bw + w2[k], bw).astype(F)
    safe = np.where(bw > 0.01, bw, F(1))
    vacc = (best / safe[..., None] * F(0.5)).astype(F); vwgt = np.full(fx.shape, F(0.5))
    for dy in (-1, 1):
        vy = (fy + F(dy) * psy * F(blur)).astype(F)
        ok = (vy >= 0) & (vy <= 1) & ((F(1) - _tex(depth, fx, vy)) > dinv + F(tol) * F(0.5))
        vacc = np.where(ok[..., None], vacc + _tex(color, fx, vy) * F(0.25), vacc).astype(F)
        vwgt = np.where(ok, vwgt + F(0.25), vwgt).astype(F)
    filled = np.where((bw > 0.01)[..., None], vacc / vwgt[..., None], _tex(color, fx, fy)).astype(F)
    mixed = (col * (F(1) - conf)[..., None] + filled * conf[..., None]).astype(F)
    col = np.where((conf > 0.001)[..., None], mixed, col)
    alpha = np.minimum(_ss(-0.001, 0.001, sx) * _ss(1.001, 0.999, sx), _ss(-0.001, 0.001, sy) * _ss(1.001, 0.999, sy))
    return col, alpha.astype(F), conf


_scene = dibr.synthetic_scene


@pytest.mark.parametrize("h,w,mode", [(40, 64, "Full-SBS"), (36, 80, "Half-SBS"), (48, 40, "Half-TAB")])
def test_c_oracle_equals_independent_numpy_restatement(h, w, mode):
    rgb, depth = _scene(h + w, h, w)
    vh, vw = dibr.view_shape(h, w, mode)
    kw = dict(ipd_uv=0.064, depth_ratio=3.0, convergence=0.2)
    left, right, cl, cr = dibr.eye_views(rgb, depth, display_mode=mode, return_conf=True, **kw)
    color = rgb.astype(F) / F(255)
    worst = 0.0
    for view, conf_c, eye in ((left, cl, -0.032), (right, cr, 0.032)):
        col, alpha, conf = numpy_eye(color, depth, vw, vh, F(eye), F(0.1 * 3.0), 0.2, (float(vw), float(vh)))
        assert (conf > 0.001).mean() > 0.01, "the scene must contain disoccluded pixels"
        # the vectorised restatement evaluates in a different association order in places: agreement to fp32 rounding
        assert np.abs(conf - conf_c).max() <= 1e-5
        assert np.abs(alpha - view[..., 3]).max() <= 1e-5
        worst = max(worst, float(np.abs(col - view[..., :3]).max()))
    assert worst <= 2e-5, worst


def test_known_answers():
    h, w = 24, 32
    rgb, depth = _scene(1, h, w)
    # (i) zero eye separation: no parallax, no disocclusion -> each eye view is the frame itself (texel centres are sampled to within
    # the fp32 rounding of (j + 0.5) / w * w - 0.5)
    left, right, cl, cr = dibr.eye_views(rgb, np.full((h, w), 0.5, np.float32), ipd_uv=0.0, display_mode="Full-SBS", return_conf=True)
    assert np.abs(left[..., :3] - rgb.astype(F) / F(255)).max() <= 1e-6 and np.array_equal(left, right)
    assert (left[..., 3] == 1).all() and (cl == 0).all()
    # (ii) flat depth at the screen plane (shift = shaped + convergence = 0): identity as well
    d = np.full((h, w), 0.4, np.float32)
    conv = float(0.4 * (1 + 0.35 * 0.6))
    left, right = dibr.eye_views(rgb, d, ipd_uv=0.064, depth_ratio=2.0, convergence=conv, display_mode="Full-SBS")
    assert np.abs(left[..., :3] - rgb.astype(F) / F(255)).max() <= 2e-3      # (convergence is rounded to fp32: a ~1e-8 uv shift)
    # (iii) the eyes shift in opposite directions and the packed frame has the viewer's layout
    _, depth = _scene(2, h, w)
    sbs = dibr.make_sbs_dibr_oracle(rgb, depth, "Full-SBS", depth_ratio=3.0)
    tab = dibr.make_sbs_dibr_oracle(rgb, depth, "Full-TAB", depth_ratio=3.0)
    assert sbs.shape == (h, 2 * w, 3) and tab.shape == (2 * h, w, 3)
    assert np.array_equal(sbs[:, :w], tab[:h]) and np.array_equal(sbs[:, w:], tab[h:])
    assert not np.array_equal(sbs[:, :w], sbs[:, w:])
    half = dibr.make_sbs_dibr_oracle(rgb, depth, "Half-SBS", depth_ratio=3.0)
    assert half.shape == (h, w, 3)
