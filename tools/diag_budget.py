"""Diagnostic: is test_bounded_walks_are_unbiased[mixed] failing on noise or on bias?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import scenes
from imgmetrics import rel_mse, luminance
from narvalengine_b200.engine import Context

ctx = Context(0)
b = scenes.mixed_scene()
cam = scenes.MIXED_CAMERA
def render(budget, seed, spp=8192, W=24, H=16):
    os.environ["NE_B200_TRACK_BUDGET"] = str(budget)
    ctx.upload(b)
    lin = np.zeros((H, W, 3), np.float32)
    ctx.render_frame(cam.make(W / H, ctx.lib), W, H, spp, 6, seed, 0, None, lin)
    return lin
u12 = render(10**8, 12); u13 = render(10**8, 13); a11 = render(3, 11); a12 = render(3, 12); a64 = render(64, 14)
print("floor  inf/12 vs inf/13", rel_mse(u13, u12), luminance(u13).mean() / luminance(u12).mean())
print("b3/11 vs inf/12", rel_mse(a11, u12), luminance(a11).mean() / luminance(u12).mean())
print("b3/12 vs inf/12", rel_mse(a12, u12), luminance(a12).mean() / luminance(u12).mean())
print("b3/11 vs b3/12", rel_mse(a11, a12))
print("b64/14 vs inf/12", rel_mse(a64, u12))
# per-pixel worst offenders
d = (a11 - u12) ** 2 / (u12 ** 2 + 1e-4)
idx = np.argsort(d.sum(-1).ravel())[-5:]
for i in idx:
    y, x = divmod(int(i), 24)
    print((y, x), a11[y, x], u12[y, x], u13[y, x], a12[y, x])
