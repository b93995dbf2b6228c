"""Cuts the large SRS fixture the real-SRS GPU tests use out of the reference's own SRS file: the 80-byte ptau header and
the first 2^16 + 3 G1 records of /root/reference/srs/srs00 (reference-owned bytes, copied verbatim; 4.2 MB).

    python tools/cut_srs_fixture.py            ->  tests/golden/_large/srs_2_16.bin

The output is git-ignored (history stays small) but NOT gpurun-ignored: it ships to the GPU box with the repository
snapshot, where /root/reference does not exist.  tests/conftest.py runs this automatically when the reference is present
and the fixture is missing.  The two G2 elements are already committed (tests/golden/srs_g2.bin)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/srs/srs00"
OUT = os.path.join(ROOT, "tests", "golden", "_large", "srs_2_16.bin")
N = (1 << 16) + 3


def main() -> int:
    if not os.path.exists(SRC):
        print("reference SRS not present; nothing cut", file=sys.stderr)
        return 1
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(SRC, "rb") as f:
        data = f.read(80 + 64 * N)
    assert data[:4] == b"ptau" and len(data) == 80 + 64 * N
    # must agree with the committed small head (same file, same offsets)
    with open(os.path.join(ROOT, "tests", "golden", "srs_head.bin"), "rb") as f:
        head = f.read()
    assert data[:len(head)] == head, "srs_head.bin is not a prefix of the reference SRS"
    with open(OUT, "wb") as f:
        f.write(data)
    print(f"wrote {OUT}: {len(data)} bytes, {N} G1 records")
    return 0


if __name__ == "__main__":
    sys.exit(main())
