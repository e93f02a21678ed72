"""In-memory quantisation state (replaces the reference's ./output_pt/** file protocol).

The reference passes every piece of inter-node state through files relative to the cwd: thresholds
(output_pt/input/input.{i}.scale.pt / .zero.pt), running min/max (.max_val.pt / .min_val.pt), weight
codes and scales (output_pt/weight/conv.weight.{i}.pt / .scale.pt), the residual shortcut tensor.  Here the
same items live in one registry object.  With define.COMPAT_FILES the registry still reads thresholds from,
and mirrors thresholds / running min-max to, the reference's files, so test.py's calibration tail
(test.py:186-217) and a later sim.py run keep working unchanged.
"""
import os
from typing import Dict, Optional

import torch


def _compat() -> bool:
    import define
    return bool(getattr(define, "COMPAT_FILES", True))


class QuantState:
    INPUT_DIR = "output_pt/input"

    def __init__(self):
        self.reset()

    def reset(self):
        self.scales: Dict[int, float] = {}
        self.zeros: Dict[int, int] = {}
        self.run_min: Dict[int, float] = {}
        self.run_max: Dict[int, float] = {}
        self.weight_codes: Dict[int, torch.Tensor] = {}     # int8 CPU tensors
        self.weight_scale: Dict[int, float] = {}
        self.bias: Dict[int, torch.Tensor] = {}
        self.bias_quan: Dict[int, torch.Tensor] = {}
        self.add_const: Dict[int, torch.Tensor] = {}
        self.requan: Dict[str, tuple] = {}                  # "0_1" -> (M16, n)
        self.shortcut: Optional[torch.Tensor] = None
        self.version = 0                                    # bumped whenever thresholds change

    # ---- thresholds ---------------------------------------------------------------------------------
    def _file(self, i, what):
        return os.path.join(self.INPUT_DIR, "input.{}.{}.pt".format(i, what))

    def scale(self, i: int) -> float:
        if i not in self.scales:
            if _compat() and os.path.isfile(self._file(i, "scale")):
                self.scales[i] = float(torch.load(self._file(i, "scale")))
            else:
                raise KeyError("no scale for activation domain %d: calibrate first (myQL.state / test.py)" % i)
        return self.scales[i]

    def zero(self, i: int) -> int:
        if i not in self.zeros:
            if _compat() and os.path.isfile(self._file(i, "zero")):
                self.zeros[i] = int(torch.load(self._file(i, "zero")))
            else:
                raise KeyError("no zero point for activation domain %d: calibrate first" % i)
        return self.zeros[i]

    def set_threshold(self, i: int, scale: float, zero: int, write_file: Optional[bool] = None):
        self.scales[i] = float(scale)
        self.zeros[i] = int(zero)
        self.version += 1
        if _compat() if write_file is None else write_file:
            os.makedirs(self.INPUT_DIR, exist_ok=True)
            torch.save(float(scale), self._file(i, "scale"))
            torch.save(int(zero), self._file(i, "zero"))

    def forget_thresholds(self):
        """Drop cached thresholds so the next use re-reads the files (e.g. after test.py's tail rewrote them)."""
        self.scales.clear()
        self.zeros.clear()
        self.version += 1

    # ---- running min / max (quan_func.py:184-195) ----------------------------------------------------
    def observe(self, i: int, lo: float, hi: float):
        lo, hi = float(lo), float(hi)
        first = i not in self.run_max
        if _compat() and os.path.isfile(self._file(i, "max_val")):
            # the files are authoritative in compat mode: test.py:108-113 resets them behind our back
            self.run_max[i] = float(torch.load(self._file(i, "max_val")))
            self.run_min[i] = float(torch.load(self._file(i, "min_val")))
            first = False
        changed = False
        if first:
            self.run_max[i], self.run_min[i] = hi, lo
            changed = True
        else:
            if self.run_max[i] < hi:
                self.run_max[i] = hi
                changed = True
            if self.run_min[i] > lo:
                self.run_min[i] = lo
                changed = True
        if changed and _compat():
            os.makedirs(self.INPUT_DIR, exist_ok=True)
            torch.save(self.run_max[i], self._file(i, "max_val"))
            torch.save(self.run_min[i], self._file(i, "min_val"))


state = QuantState()
