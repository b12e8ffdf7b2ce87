"""Writes prior_diffuse_b200/csrc/time_table.inc: the float32 bit patterns of DiffUNet1's sinusoid table
(model/diff3.py:89-95, TimeEmbedding._build_embedding) exactly as torch computes it -- the C packer ships them because
one ulp of the float32 argument (up to 4.9e5) moves sin / cos by percents.   python tests/golden/make_time_table.py"""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def table_bits():
    arg = (torch.arange(50).unsqueeze(1) * 10.0 ** (torch.arange(64).unsqueeze(0) * 4.0 / 63.0))
    table = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1).contiguous()
    return table.numpy().view(np.uint32).reshape(-1)


if __name__ == "__main__":
    bits = table_bits()
    with open(os.path.join(ROOT, "prior_diffuse_b200", "csrc", "time_table.inc"), "w") as f:
        for i in range(0, bits.size, 8):
            f.write(", ".join("0x%08xu" % int(b) for b in bits[i:i + 8]) + ",\n")
    print("wrote", bits.size, "words")
