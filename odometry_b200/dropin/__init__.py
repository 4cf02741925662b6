"""Drop-in C++ classes (ImagePyramid / DepthPyramid / DepthEstimator / LevenbergMarquardtOptimizer
and the free functions of image_processing_global.h) over the C-ABI of libodo_b200.so."""
