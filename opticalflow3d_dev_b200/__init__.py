"""
opticalflow3d_dev_b200 -- B200-native (sm_100a) dense Lucas-Kanade optical flow for 2D/3D
microscopy time-lapses, a drop-in for the `calc_flow` module of ScientistRachel/OpticalFlow3D_dev.

    from opticalflow3d_dev_b200.calc_flow import calc_flow3D, calc_flow2D, process_flow

Importing the package does not touch the GPU; the first call loads libof3d.so and fails loudly
if the library or a CUDA device is missing (there is no CPU fallback).
"""
from .calc_flow import calc_flow2D, calc_flow3D, process_flow  # noqa: F401
from .analysis import masked_flow, reliability_threshold  # noqa: F401

__version__ = '0.1.0'
