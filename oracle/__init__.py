"""CPU oracle for the multi-view fusion hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy fp32 restatement of the reference's arithmetic for the
path  unproject -> fuse across views -> project -> PyramidROIAlign -> per-class NMS
(reference: mrcnn/model_multi.py, mrcnn/recurrent.py, Notebook/projection.py).

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the *checker* or the *timed CPU
baseline*, never as the product.  Nothing under ``mulit_view_object_detection_b200/``
imports it, and the product raises when its CUDA library is missing.

Parity pinning status
---------------------
The reference ships no tests, fixtures or golden vectors and needs
TensorFlow 1.x + Keras 2.x, which are not installable here.  What pins this oracle:

* ``tests/golden/make_golden.py`` executes the reference's OWN Python source for
  ``unproj_feat``, ``proj_grid``/``nearest3``, ``apply_box_deltas_graph``,
  ``clip_boxes_graph``, ``refine_detections_graph``, ``PyramidROIAlign.call``,
  ``ProposalLayer.call`` and ``ConvLSTMCell.call`` (imported from /root/reference in the
  build container) over a small eager NumPy stand-in for the ``tf`` namespace, and stores
  inputs/outputs as fixtures under ``tests/golden/``.  That pins every reference-authored
  decision (operand order, transposes, meshgrid order, index layout, gate order ...).
  ``grid_reas`` ('add', 'ident', 'conv3d') and ``depth_sampling`` (both branches) -- Keras graph builders -- are executed
  with functional stand-ins for the layers they instantiate, weights looked up by the reference's layer names: that pins the
  wiring (channel orders, ReLU placement, skip-concat order, names), not the Keras kernels inside the layers.
* the pure-NumPy helpers of the reference (``mrcnn/utils.py``: ``compute_iou``,
  ``non_max_suppression``, ``apply_box_deltas``, ``vec2rot``, ``quat2rot``) run unmodified.
* The third-party TensorFlow kernels themselves (``gather_nd`` out-of-range behaviour,
  ``crop_and_resize``, ``non_max_suppression``, ``range``/``linspace`` fill order, matmul
  contraction order) are restated from their published algorithms; TensorFlow's version is
  unpinned by the reference (``assert tf >= 1.3``).  For those ops parity with real TF bits
  is *unpinned* -- the evaluation order in SURVEY.md Appendix A is the definition.

All arithmetic is fp32 with every ``*``/``+``/``/`` individually rounded (no FMA) and dot
products evaluated left-to-right in ascending k.
"""
from .geometry import (tf1_range, tf1_linspace, matmul_seq, unproj_matrices,
                       grid_centres, proj_constants)
from .unproject import unproj_feat, unproj_feat_notebook, unproject_coords
from .fusion import (grid_reas, fuse_views, batch_norm_affine, ident_fuse, convlstm,
                     convlstm_cell_step, channel_mean, unet_fuse, depth_sampling_conv3d,
                     conv3d_strided_same, conv3d_transpose_same, fusion_neck)
from .projection import proj_grid, depth_sampling, project_indices
from .roi_align import pyramid_roi_align, roi_levels, crop_and_resize
from .detection import (apply_box_deltas, clip_boxes, iou_tf, non_max_suppression,
                        refine_detections, detection_layer, proposal_layer, norm_boxes)
from . import model
