"""Image metrics of SURVEY.md 8(d): rel-MSE on linear buffers (eps = 1e-4) and Rec.709 luminance."""
import numpy as np


def rel_mse(a, b, eps=1e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.mean((a - b) ** 2 / (b * b + eps)))


def luminance(img):
    img = np.asarray(img, np.float64)
    return 0.2126 * img[..., 0] + 0.7152 * img[..., 1] + 0.0722 * img[..., 2]
