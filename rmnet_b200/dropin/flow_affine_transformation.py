"""Top-level module named like the reference's NumPy C extension (`import flow_affine_transformation`,
utils/data_transforms.py:18).  update_optical_flow(of, M1, M2) -> ndarray   (flow_affine_transformation.cpp:87-90)."""
from rmnet_b200.ops import update_optical_flow  # noqa: F401
