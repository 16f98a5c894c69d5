"""Top-level alias so that the reference's `from calc_flow import process_flow` (example_processing_script.ipynb:10)
resolves to the B200 implementation when the repository root is on sys.path."""
from opticalflow3d_dev_b200.calc_flow import calc_flow2D, calc_flow3D, process_flow  # noqa: F401

__all__ = ['calc_flow2D', 'calc_flow3D', 'process_flow']
