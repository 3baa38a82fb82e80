"""CPU: Keras layer names -> parameter dictionaries (weights_io.py) and the image_meta vector, checked by feeding the mapped
parameters to the oracle's fusion neck (mrcnn/model_multi.py:2382-2410) and against the reference's own
compose/parse_image_meta when /root/reference is present."""
import os

import numpy as np
import pytest

import oracle
from helpers import small_cfg, scene
from mulit_view_object_detection_b200 import weights_io as wio

LEVELS = (4, 5)


def _bn(rng, n):
    return [rng.uniform(0.8, 1.2, n).astype(np.float32), rng.normal(0, 0.05, n).astype(np.float32),
            rng.normal(0, 0.05, n).astype(np.float32), rng.uniform(0.7, 1.3, n).astype(np.float32)]


def _named(mode, rng, V, C, F, S):
    named = {}
    for lvl in LEVELS:
        scope, depth = "grid_reas_P%d" % lvl, "grid_reas_depth_PG%d" % lvl
        if mode == "conv3d":
            shapes = {"_1": (3, 3, 3, V * C, 2 * F), "_2": (3, 3, 3, 2 * F, 4 * F), "_deconv_1": (3, 3, 3, 2 * F, 4 * F),
                      "_deconv_2": (3, 3, 3, F, 4 * F)}
            for suf, shp in shapes.items():
                nout = shp[3] if "deconv" in suf else shp[4]
                named[scope + "_3D_conv" + suf] = [(rng.standard_normal(shp) * 0.05).astype(np.float32), rng.normal(0, 0.1, nout).astype(np.float32)]
                named[scope + "_batch_norm" + (suf if "deconv" not in suf else suf[1:])] = _bn(rng, nout)
            for i, (cin, cout) in enumerate(((F * S, 8), (8, F)), 1):
                named[depth + "_DepthwiseConv_%d" % i] = [rng.uniform(0.5, 1.5, (1, 1, cin, 1)).astype(np.float32), rng.normal(0, 0.1, cin).astype(np.float32)]
                named[depth + "2DConv_%d" % i] = [(rng.standard_normal((1, 1, cin, cout)) * 0.1).astype(np.float32), rng.normal(0, 0.1, cout).astype(np.float32)]
                named[depth + "bn_%d" % i] = _bn(rng, cout)
            continue
        named[scope + "_batch_norm"] = _bn(rng, F)
        if mode == "ident":
            named[scope + "ident_conv"] = [(rng.standard_normal((1, 1, 1, V * C, F)) * 0.1).astype(np.float32), rng.normal(0, 0.1, F).astype(np.float32)]
        if mode == "lstm3d":
            named[scope + "_convlstm3d"] = [(rng.standard_normal((3, 3, 3, C + F, 4 * F)) * 0.05).astype(np.float32), np.zeros(4 * F, np.float32)]
        named[depth + "2DConv"] = [rng.normal(0.1, 0.3, (1, 1, S, 1)).astype(np.float32), np.array([0.05], np.float32)]
        named[depth + "bn_deconv"] = _bn(rng, 1)
    return named


@pytest.mark.parametrize("mode", ["add", "ident", "lstm3d", "conv3d"])
def test_keras_names_map_to_neck_params(mode, tmp_path):
    rng = np.random.default_rng(3)
    V, C, F, S = 2, 4, 4, 3
    cfg = small_cfg(GRID_REAS=mode, NUM_VIEWS=V, nvox=4, nvox_z=4, samples=S, TOP_DOWN_PYRAMID_SIZE=F,
                    IMAGE_SHAPE=np.array([64, 64, 3]), VANILLA=True)
    named = _named(mode, rng, V, C, F, S)
    path = os.path.join(str(tmp_path), "w.npz")
    wio.write_npz(path, named)
    back = wio.read_npz(path)
    assert set(back) == set(named) and all(np.array_equal(a, b) for k in named for a, b in zip(named[k], back[k]))
    params = wio.fusion_params_from_keras(back, cfg, levels=LEVELS)
    assert set(params) == {"grid_reas_P4", "grid_reas_P5", "grid_reas_depth_PG4", "grid_reas_depth_PG5"}
    fmaps = []
    for lvl in LEVELS:
        f, Rcam, Kmat = scene(cfg, 1, V, 64 >> lvl, 64 >> lvl, C, seed=1, image_hw=(64, 64))
        fmaps.append(f)
    outs = oracle.fusion_neck(fmaps, Rcam, Kmat, cfg, params, levels=LEVELS)
    assert [o.shape for o in outs] == [(1, 4, 4, F), (1, 2, 2, F)]
    assert all(np.isfinite(o).all() for o in outs)
    if mode == "ident":
        assert params["grid_reas_P4"]["weight"].shape == (V * C, F)
    if mode == "conv3d":
        assert params["grid_reas_depth_PG4"]["conv1"]["W"].shape == (F * S, 8)


def test_missing_layer_is_reported():
    cfg = small_cfg(GRID_REAS="ident")
    with pytest.raises(KeyError, match="ident_conv"):
        wio.fusion_params_from_keras({}, cfg, levels=(4,))


def test_image_meta_roundtrip_and_reference():
    meta = wio.compose_image_meta(7, (480, 640, 3), (640, 640, 3), (80, 0, 560, 640), 1.0, np.arange(5) % 2)
    assert meta.shape == (12 + 5,)
    p = wio.parse_image_meta(meta[None])
    assert p["image_id"][0] == 7 and p["window"][0].tolist() == [80, 0, 560, 640] and p["scale"].dtype == np.float32
    assert p["original_image_shape"][0].tolist() == [480, 640, 3] and p["active_class_ids"][0].tolist() == [0, 1, 0, 1, 0]
