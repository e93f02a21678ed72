"""Hardware test-vector export: ./output_pt/** -> ./output_txt/** hex text (mirror of the reference's output.py:41-245;
run as `python output.py` from the directory holding ./output_pt, exactly like the reference, README.md:12).

The reference walks every element with four nested Python loops and a `.item()` call (5.7 s for one 23x34 frame);
here each block of rows is formatted at once with numpy (bit-identical files, checked against the reference's own
output in tests/test_output_export.py).  Formats (all values two's-complement hex, 32-wide tiles):
  input/input.{0..5}.txt            activation codes, QUAN_BIT; tiles shrink by the conv halo per layer (:41-118)
  bias/param_buf.txt                per layer and channel: bias const (BIAS_BIT) | M16 (REQUAN_BIT) | Mres (:120-139)
  pe_out/pe_output{i}_{p}.txt       per-PE partial sums, PE_ACC_BIT (:141-186)
  pe_add/pe_add_output{i}.txt       adder-tree outputs, PE_ADD_BIT (:188-226)
  requan_shift_n/requan_shift_n.txt the six shift counts (:228-245)
"""
import math
import os

import numpy as np
import torch

from define import PE_ACC_BIT, PE_ADD_BIT, QUAN_BIT, REQUAN_BIT, REQUAN_N_MAX, BIAS_BIT

tile_width = 32
_HEX = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)


def hex_digits(bit_width):
    """Digits the reference prints for a bit width (3..8 nibbles honoured, anything else -> 2)."""
    n = math.ceil(bit_width / 4)
    return n if n in (3, 4, 5, 6, 7, 8) else 2


def float_to_hex(item, bit_width):
    """signed integer value -> two's-complement hex string (reference output.py:13-39)."""
    value = int(item)
    if value < 0:
        value += int(2 ** bit_width)
    return '{:0{w}x}'.format(value, w=hex_digits(bit_width))


def hex_rows(block, bit_width):
    """block: 2-D array of integer values -> one hex string per row (each value hex_digits(bit_width) wide)."""
    v = np.asarray(block).astype(np.int64)
    v = np.where(v < 0, v + (1 << int(bit_width)), v).astype(np.uint64)
    nd = hex_digits(bit_width)
    if v.size and int(v.max()) >= 16 ** nd:
        return ["".join(float_to_hex(x, bit_width) for x in row) for row in np.asarray(block)]   # wider than the field
    shifts = np.arange(nd - 1, -1, -1, dtype=np.uint64) * np.uint64(4)
    chars = _HEX[((v[..., None] >> shifts) & np.uint64(15)).astype(np.intp)]
    return [bytes(r).decode("ascii") for r in chars.reshape(v.shape[0], -1)]


def _load(path):
    t = torch.load(path, map_location="cpu")
    return t.detach().cpu().numpy() if hasattr(t, "detach") else t


def _open(folder, name):
    os.makedirs(folder, exist_ok=True)
    return open(os.path.join(folder, name), "w")


def export_inputs(src="output_pt/input", dst="output_txt/input"):
    """Activation codes of the six domains, tiled 32 wide; the first tile of every row/column of tiles shrinks by the
    accumulated conv halo (5,3,3,3,5 -> 2+1+1+1+2) so that tiles stay aligned across layers (reference :41-118)."""
    conv_kernel_size = [0, 5, 3, 3, 3, 5]
    h_overlap = w_overlap = tile_width
    for layer_id in range(6):
        data = _load(os.path.join(src, "input.{}.pt".format(layer_id)))
        _, channel, height, width = data.shape
        expanded_height = ((height // tile_width) + 1) * tile_width
        expanded_width = ((width // tile_width) + 1) * tile_width
        expand = np.zeros((channel, height, expanded_width), dtype=np.int64)
        expand[:, :, :width] = np.rint(data[0]).astype(np.int64)
        width_block_num = expanded_width // tile_width
        height_block_num = expanded_height // tile_width
        h_overlap -= conv_kernel_size[layer_id] // 2
        w_overlap -= conv_kernel_size[layer_id] // 2
        with _open(dst, "input.{}.txt".format(layer_id)) as f:
            out = []
            block_h_start = 0
            for hb in range(height_block_num):
                block_w_start = 0
                current_height = h_overlap if hb == 0 else tile_width
                for wb in range(width_block_num):
                    current_width = w_overlap if wb == 0 else tile_width
                    if hb == height_block_num - 1:
                        current_height = height - block_h_start
                    out.append("{:02x}".format(int(current_height)))
                    out.append("{:02x}".format(int(channel)))
                    pad = "00" * (32 - current_width)
                    for c in range(channel):
                        out.append("{:02x}".format(c))
                        rows = hex_rows(expand[c, block_h_start:block_h_start + max(current_height, 0),
                                               block_w_start:block_w_start + current_width], QUAN_BIT)
                        out.extend(r + pad for r in rows)
                    block_w_start += current_width
                block_h_start += current_height
            f.write("\n".join(out) + "\n")


def export_bias(src="output_pt", dst="output_txt/bias"):
    with _open(dst, "param_buf.txt") as f:
        f.write(float_to_hex(5, 8) + "\n")
        requan_res = float_to_hex(_load(os.path.join(src, "requan_factor/requan_res.pt")), REQUAN_BIT)
        for layer_id in range(5):
            bias_quan = _load(os.path.join(src, "bias/conv.bias.quan{}.pt".format(layer_id)))
            factor = _load(os.path.join(src, "requan_factor/requan_{}_{}.pt".format(layer_id, layer_id + 1)))
            chnl = bias_quan.shape[1]
            f.write(float_to_hex(chnl, 8) + "\n")
            for ci in range(chnl):
                f.write(float_to_hex(bias_quan[0, ci, 0, 0], BIAS_BIT) + float_to_hex(factor, REQUAN_BIT) + requan_res + "\n")


def _export_tiled(data, bit_width, f):
    """[C, H, W] values in 32 x 32 tiles: per tile a header (rows, channels), per channel its index and the rows of the
    tile that exist, every row 32 values wide, zero padded on the right (reference :141-226)."""
    channel, height, width = data.shape
    expanded_height = height if height % tile_width == 0 else ((height // tile_width) + 1) * tile_width
    expanded_width = width if width % tile_width == 0 else ((width // tile_width) + 1) * tile_width
    expand = np.zeros((channel, height, expanded_width), dtype=np.int64)
    expand[:, :, :width] = np.rint(data).astype(np.int64)
    out = []
    for hb in range(expanded_height // tile_width):
        for wb in range(expanded_width // tile_width):
            h0, w0 = hb * tile_width, wb * tile_width
            lines = height - h0 if hb == expanded_height // tile_width - 1 else 32
            out.append("{:02x}".format(int(lines)))
            out.append("{:02x}".format(int(channel)))
            for c in range(channel):
                out.append("{:02x}".format(c))
                out.extend(hex_rows(expand[c, h0:min(h0 + 32, height), w0:w0 + 32], bit_width))
    f.write("\n".join(out) + "\n")


def export_pe_out(src="output_pt/pe_out", dst="output_txt/pe_out"):
    for layer_id in range(5):
        for pe_id in range(4):
            data = _load(os.path.join(src, "pe_output{}_{}.pt".format(layer_id, pe_id)))       # saved 3-D (reference :146)
            with _open(dst, "pe_output{}_{}.txt".format(layer_id, pe_id)) as f:
                _export_tiled(data, PE_ACC_BIT, f)


def export_pe_add(src="output_pt/pe_add", dst="output_txt/pe_add"):
    for layer_id in range(5):
        data = _load(os.path.join(src, "pe_add_output{}.pt".format(layer_id)))
        with _open(dst, "pe_add_output{}.txt".format(layer_id)) as f:
            _export_tiled(data[0], PE_ADD_BIT, f)


def export_requan_shift_n(src="output_pt/requan_factor", dst="output_txt/requan_shift_n"):
    with _open(dst, "requan_shift_n.txt") as f:
        bits = math.log2(REQUAN_N_MAX)
        for layer_id in range(5):
            f.write(float_to_hex(_load(os.path.join(src, "n_{}_{}.pt".format(layer_id, layer_id + 1))), bits) + "\n")
        f.write(float_to_hex(_load(os.path.join(src, "n_res.pt")), bits))


def export_all():
    export_inputs()
    export_bias()
    export_pe_out()
    export_pe_add()
    export_requan_shift_n()


if __name__ == "__main__":
    export_all()
