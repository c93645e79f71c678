"""B200-native (sm_100a) replacement for the native hot path of Bundle-Adjusting-Gaussian-Splatting / Self-Cali-GS.

The directory holding this file is meant to be put on ``sys.path`` (or ``PYTHONPATH``): it then provides the two
packages the reference imports -- ``diff_gaussian_rasterization`` and ``simple_knn`` -- backed by
``libbags_b200.so`` (C ABI in ``include/bags_b200.h``).  Importing this package does the path insertion.
"""
import os
import sys

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
if PACKAGE_DIR not in sys.path:
    sys.path.insert(0, PACKAGE_DIR)
