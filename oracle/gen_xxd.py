#!/usr/bin/env python3
"""Stand-in for `xxd -i parametersDefault` (xxd is not installed in this image).

Writes the C array the reference embeds at compile time (reference source/Makefile:149-150,
consumed by source/Parameters.cpp:317 and source/STAR.cpp:34).  Output goes to oracle/_ref/gen/.
"""
import sys

src, dst = sys.argv[1], sys.argv[2]
data = open(src, "rb").read()
with open(dst, "w") as f:
    f.write("unsigned char parametersDefault[] = {\n")
    for i in range(0, len(data), 12):
        f.write("  " + ", ".join("0x%02x" % b for b in data[i:i + 12]) + ",\n")
    f.write("};\nunsigned int parametersDefault_len = %d;\n" % len(data))
