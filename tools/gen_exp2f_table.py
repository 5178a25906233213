#!/usr/bin/env python
"""The 32-entry table of the expf restatement (oracle/tf_oracle.c, svt-av1_b200/csrc/tf.cu):
T[i] = bits(2^(i/32), correctly rounded to double) - (i << 47).  Needs mpmath (200-bit arithmetic)."""
import struct

import mpmath as mp

mp.mp.prec = 200
for i in range(32):
    bits = struct.unpack("<Q", struct.pack("<d", float(mp.power(2, mp.mpf(i) / 32))))[0]
    print("0x%016xull," % (bits - (i << 47)))
