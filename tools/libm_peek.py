"""Reads doubles out of this image's libm.so.6 by virtual address (development aid for oracle/glibc_trig.c:
the restatement of glibc 2.39's x86_64 FMA-variant sin / cos / atan2 was checked against the library's own
constants and tables; nothing here is used at run time)."""
import struct
import sys

LIBM = "/usr/lib/x86_64-linux-gnu/libm.so.6"


def sections(data):
    shoff = struct.unpack_from("<Q", data, 0x28)[0]
    shentsize, shnum = struct.unpack_from("<HH", data, 0x3A)
    out = []
    for i in range(shnum):
        off = shoff + i * shentsize
        _, typ, flags, addr, offset, size = struct.unpack_from("<IIQQQQ", data, off)
        out.append((addr, offset, size, typ))
    return out


def read_doubles(vaddr, count, data=None):
    data = data or open(LIBM, "rb").read()
    for addr, offset, size, typ in sections(data):
        if typ != 8 and addr <= vaddr < addr + size:      # not NOBITS
            return struct.unpack_from("<%dd" % count, data, offset + (vaddr - addr))
    raise ValueError("address not mapped")


if __name__ == "__main__":
    v = int(sys.argv[1], 16)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for k, d in enumerate(read_doubles(v, n)):
        print(hex(v + 8 * k), repr(d), d.hex())
