"""Shared helpers of the test-suite (synthetic cases, comparison utilities)."""
import numpy as np

from mulit_view_object_detection_b200.config import FusionConfig
from mulit_view_object_detection_b200 import synthetic as syn

RTOL = 1e-5          # north_star: within 1e-5 relative (fp32) for fused features and ROIAlign crops
ATOL = 1e-6          # absolute floor for values that are exactly 0 in one implementation


def small_cfg(**kw):
    base = dict(nvox=16, nvox_z=16, samples=8, NUM_VIEWS=3)
    base.update(kw)
    return FusionConfig(**base)


def scene(cfg, B=1, V=3, fh=40, fw=40, C=32, seed=0, image_hw=None):
    return syn.make_scene(cfg, B, V, fh, fw, C, seed, image_hw)


def to_dev(*arrays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def random_bn(rng, C):
    return (rng.uniform(0.5, 1.5, C).astype(np.float32), rng.normal(0, 0.2, C).astype(np.float32),
            rng.normal(0, 0.2, C).astype(np.float32), rng.uniform(0.5, 1.5, C).astype(np.float32))
