"""Seeded input frames and the case list shared by make_image_golden.py (which runs the reference's own
PrismaticImageProcessor on them) and the tests (which re-create the same frames and compare digests)."""
import numpy as np

FUSED = dict(use_fused_vision_backbone=True, input_sizes=[(3, 224, 224), (3, 224, 224)], interpolations=["bicubic", "bicubic"],
             means=[(0.485, 0.456, 0.406), (0.5, 0.5, 0.5)], stds=[(0.229, 0.224, 0.225), (0.5, 0.5, 0.5)])
SIGLIP = dict(use_fused_vision_backbone=False, input_sizes=[(3, 224, 224)], interpolations=["bicubic"],
              means=[(0.5, 0.5, 0.5)], stds=[(0.5, 0.5, 0.5)])

# (name, strategy, H, W, seed, processor kwargs key)
CASES = [
    ("libero_256_naive", "resize-naive", 256, 256, 1, "fused"),
    ("identity_224_naive", "resize-naive", 224, 224, 2, "fused"),
    ("vga_naive", "resize-naive", 480, 640, 3, "fused"),
    ("upsample_naive", "resize-naive", 100, 130, 4, "fused"),
    ("landscape_crop", "resize-crop", 200, 300, 5, "fused"),
    ("portrait_crop", "resize-crop", 300, 200, 6, "fused"),
    ("half_to_even_crop", "resize-crop", 257, 255, 7, "fused"),
    ("hd_crop", "resize-crop", 360, 640, 8, "fused"),
    ("odd_letterbox", "letterbox", 201, 300, 9, "fused"),
    ("square_letterbox", "letterbox", 256, 256, 10, "fused"),
    ("portrait_letterbox", "letterbox", 320, 180, 11, "fused"),
    ("libero_256_naive_siglip", "resize-naive", 256, 256, 12, "siglip"),
    ("landscape_letterbox_siglip", "letterbox", 240, 427, 13, "siglip"),
]
KWARGS = {"fused": FUSED, "siglip": SIGLIP}


def frame(h: int, w: int, seed: int) -> np.ndarray:
    """A camera-like uint8 [h, w, 3] frame: smooth gradients, hard edges (ringing / clipping of the bicubic filter) and noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([255 * xx / max(w - 1, 1), 255 * yy / max(h - 1, 1), 127 + 127 * np.sin(xx / 7.0 + seed) * np.cos(yy / 11.0)], -1)
    for _ in range(6):                                   # saturated rectangles: overshoot must clip at 0 / 255
        y0, x0 = rng.integers(0, h), rng.integers(0, w)
        y1, x1 = min(h, y0 + rng.integers(3, max(4, h // 3))), min(w, x0 + rng.integers(3, max(4, w // 3)))
        img[y0:y1, x0:x1] = rng.choice([0, 255], size=3)
    img += rng.normal(0, 12, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
