"""Test stub: sim.py / test.py `import cv2` but only use it in commented-out code."""
