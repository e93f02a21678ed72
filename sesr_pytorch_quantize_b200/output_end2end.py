"""End-to-end test vectors: the network input (input.0) and output (input.5) codes as 4-values-per-line hex text
(mirror of the reference's output_end2end.py:37-101; run as a script from the directory holding ./output_pt).

The reference's quirks are kept because its RTL test bench consumes exactly this text: block b of 32 rows starts at row
b (not 32*b, :66), and the "4 values per line" counter is not reset at the end of a row whose length is not a multiple
of 4 (:72-76)."""
import os

import numpy as np

from define import QUAN_BIT
from output import _load, _open, hex_rows, tile_width


def export_end2end(layer_id, src="output_pt/input", dst="output_txt/input"):
    data = _load(os.path.join(src, "input.{}.pt".format(layer_id)))
    _, channel, height, width = data.shape
    expanded_height = height if height % tile_width == 0 else ((height // tile_width) + 1) * tile_width
    expand = np.zeros((channel, expanded_height, width), dtype=np.int64)
    expand[:, :height] = np.rint(data[0]).astype(np.int64)
    out = []
    for hb in range(expanded_height // tile_width):
        out.append("{:02x}".format(hb))
        for c in range(channel):
            out.append("{:02x}".format(c))
            line_num = 0
            cur = ""
            for bh in range(tile_width):
                vals = hex_rows(expand[c, hb + bh:hb + bh + 1, :].reshape(-1, 1), QUAN_BIT)
                for v in vals:
                    cur += v
                    line_num += 1
                    if line_num == 4:
                        out.append(cur)
                        cur = ""
                        line_num = 0
                if line_num != 0:                         # reference :75-76: newline, counter NOT reset
                    out.append(cur)
                    cur = ""
    with _open(dst, "input.{}.txt".format(layer_id)) as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    export_end2end(0)
    export_end2end(5)
