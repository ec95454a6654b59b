"""Top-level module named like the reference's compiled CUDA extension (`import reg_att_map_generator`,
extensions/reg_att_map_generator/__init__.py:11): put rmnet_b200/dropin on sys.path ahead of the reference build
and the reference's own wrapper class works unchanged.  forward(mask, prob_threshold, n_pts_threshold,
n_bbox_loose_pixels) -> [att_map, bboxes]   (reg_att_map_generator_cuda.cpp:26-38)."""
from rmnet_b200.ops import reg_att_map_forward as _fwd


def forward(mask, prob_threshold, n_pts_threshold, n_bbox_loose_pixels):
    return _fwd(mask, prob_threshold, n_pts_threshold, n_bbox_loose_pixels)
