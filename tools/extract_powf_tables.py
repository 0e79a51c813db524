"""Dump the constants glibc's powf (FMA build, __powf_fma) uses, straight from the libm.so.6 of this image.

`np.float32 ** .5` on a NumPy scalar calls libm powf(x, 0.5f) (sylber/utils/segment_utils.py:69 when both
cossim arguments are 1-D).  To reproduce the reference's segment decisions bit for bit on the GPU, the CUDA
segmentation kernel replays that powf in fp64 (sylber_b200/csrc/powf_tables.cuh).  This script documents where
the numbers come from: the addresses below were read off `objdump -d libm.so.6` (glibc 2.39-0ubuntu8.5,
function at 0x7df50 selected by the powf ifunc when the CPU has FMA+AVX2):
    0xb7f80  __powf_log2_data.tab[16] {invc, logc}     0xb8080  .poly[5]  (log2 polynomial, scaled by 32)
    0xb7be0  __exp2f_data.tab[32]                       0xb7ce0  shift_scaled, 0xb7ce8 poly_scaled[3]
Run:  python tools/extract_powf_tables.py [--emit]      (--emit prints the CUDA header body)
The parity test tests/test_powf_emulation.py checks the emulation against the live libm on the box.
"""
import struct
import sys

LIBM = "/lib/x86_64-linux-gnu/libm.so.6"


def vaddr_reader(path):
    data = open(path, "rb").read()
    assert data[:4] == b"\x7fELF"
    e_phoff, = struct.unpack_from("<Q", data, 0x20)
    e_phentsize, e_phnum = struct.unpack_from("<HH", data, 0x36)
    segs = []
    for i in range(e_phnum):
        p_type, p_flags, p_offset, p_vaddr, p_paddr, p_filesz, p_memsz, p_align = struct.unpack_from(
            "<IIQQQQQQ", data, e_phoff + i * e_phentsize)
        if p_type == 1:
            segs.append((p_vaddr, p_offset, p_filesz))

    def read(vaddr, n):
        for va, off, sz in segs:
            if va <= vaddr < va + sz:
                return data[off + vaddr - va: off + vaddr - va + n]
        raise ValueError(hex(vaddr))
    return read


def main():
    rd = vaddr_reader(LIBM)
    log_tab = struct.unpack("<32d", rd(0xB7F80, 256))
    log_poly = struct.unpack("<5d", rd(0xB8080, 40))
    exp_tab = struct.unpack("<32Q", rd(0xB7BE0, 256))
    shift_scaled, c0, c1, c2 = struct.unpack("<4d", rd(0xB7CE0, 32))
    minus_one, = struct.unpack("<d", rd(0x99360, 8))
    one, = struct.unpack("<d", rd(0x98E18, 8))
    assert minus_one == -1.0 and one == 1.0, (minus_one, one)
    assert log_tab[2 * 8] == 1.0 or 1.0 in log_tab[0::2], "log2 table should contain invc == 1"
    assert exp_tab[0] == 0x3FF0000000000000
    if "--emit" in sys.argv:
        print("__device__ const double kPowfLogTab[32] = {")
        for i in range(16):
            print(f"    {log_tab[2*i].hex()}, {log_tab[2*i+1].hex()},")
        print("};")
        print("__device__ const double kPowfLogPoly[5] = {" + ", ".join(x.hex() for x in log_poly) + "};")
        print("__device__ const unsigned long long kExp2fTab[32] = {")
        for i in range(0, 32, 4):
            print("    " + ", ".join(f"0x{v:016x}ull" for v in exp_tab[i:i+4]) + ",")
        print("};")
        print(f"constexpr double kExp2fShiftScaled = {shift_scaled.hex()};")
        print("constexpr double kExp2fPolyScaled[3] = {" + ", ".join(x.hex() for x in (c0, c1, c2)) + "};")
    else:
        print("log2 tab", [(a.hex(), b.hex()) for a, b in zip(log_tab[0::2], log_tab[1::2])])
        print("log2 poly", [x.hex() for x in log_poly])
        print("exp2 tab", [hex(v) for v in exp_tab])
        print("shift_scaled", shift_scaled.hex(), "poly_scaled", c0.hex(), c1.hex(), c2.hex())


if __name__ == "__main__":
    main()
