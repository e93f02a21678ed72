"""Test stub for scikit-image (not in this image): only skimage.metrics, as far as test.py / sim.py use it."""
