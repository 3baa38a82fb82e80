"""Configuration object for the fusion hot path.

Mirrors the attribute NAMES the reference's layers read from its ``Config`` subclasses
(mrcnn/config.py:17-236 and samples/interior/interior_multi.py:370-421), so a reference
``Config`` instance can be passed to the layers in :mod:`.layers` unchanged -- the layers are
duck-typed on these attributes.  Only the attributes the hot path reads exist here.
"""
import numpy as np


class FusionConfig:
    # batch / image geometry (config.py:208-224)
    GPU_COUNT = 1
    IMAGES_PER_GPU = 1
    IMAGE_SHAPE = np.array([640, 640, 3])
    NUM_VIEWS = 2
    NUM_CLASSES = 23
    TOP_DOWN_PYRAMID_SIZE = 256
    TRAIN_BN = False

    # voxel grid (interior_multi.py:379-389)
    GRID_REAS = "add"
    VANILLA = False          # interior_multi.py:393,421: PG2 / PG3 are replaced by zeros (model_multi.py:2406-2410)
    nvox = 40
    nvox_z = 40
    vmin = -2.5
    vmax = 2.5
    vmin_z = 1.0
    vmax_z = 10.0
    samples = 20

    # ROIAlign / detection / proposals (config.py:89-175)
    POOL_SIZE = 7
    MASK_POOL_SIZE = 14
    DETECTION_MAX_INSTANCES = 100
    DETECTION_MIN_CONFIDENCE = 0.7
    DETECTION_NMS_THRESHOLD = 0.3
    BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    RPN_NMS_THRESHOLD = 0.7
    PRE_NMS_LIMIT = 6000
    POST_NMS_ROIS_INFERENCE = 1000

    def __init__(self, **overrides):
        for k, v in overrides.items():
            setattr(self, k, v)
        self.BATCH_SIZE = self.IMAGES_PER_GPU * self.GPU_COUNT
        # derived exactly as the reference configs do (interior_multi.py:385-386)
        if "vsize" not in overrides:
            self.vsize = float(self.vmax - self.vmin) / self.nvox
        if "vsize_z" not in overrides:
            self.vsize_z = float(self.vmax_z - self.vmin_z) / self.nvox_z
        self.IMAGE_SHAPE = np.asarray(self.IMAGE_SHAPE)
        self.IMAGE_META_SIZE = 1 + 3 + 3 + 4 + 1 + self.NUM_CLASSES

    def display(self):
        for a in sorted(dir(self)):
            if not a.startswith("__") and not callable(getattr(self, a)):
                print("{:30} {}".format(a, getattr(self, a)))
