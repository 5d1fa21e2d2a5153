"""Builds tests/golden/sign_input.bin.gz from the reference's own golden file.

Run HERE (container with /root/reference); the output is committed so that the
GPU box (no /root/reference) can use it.

Source: /root/reference/src/sign/eddsa/testdata/sign.input.gz — 1024 lines
``seed||pk : pk : msg : sig||msg :`` in hex, message length 0..1023, consumed by the
reference's ``tests/sign/eddsa.rs:37-94`` (test_golden).

Output record layout (little-endian), repeated 1024 times:
    seed[32] pk[32] sig[64] msg_len:u32 msg[msg_len]
"""
import gzip
import struct
import sys

SRC = "/root/reference/src/sign/eddsa/testdata/sign.input.gz"
DST = __file__.rsplit("/", 1)[0] + "/sign_input.bin.gz"


def main():
    out = bytearray()
    n = 0
    with gzip.open(SRC, "rt") as f:
        for line in f:
            parts = line.strip().split(":")
            if len(parts) < 4:
                continue
            skpk = bytes.fromhex(parts[0])
            pk = bytes.fromhex(parts[1])
            msg = bytes.fromhex(parts[2])
            sigmsg = bytes.fromhex(parts[3])
            assert len(skpk) == 64 and skpk[32:] == pk and len(pk) == 32
            assert sigmsg[64:] == msg
            out += skpk[:32] + pk + sigmsg[:64] + struct.pack("<I", len(msg)) + msg
            n += 1
    assert n == 1024, n
    with gzip.GzipFile(DST, "wb", mtime=0) as g:
        g.write(bytes(out))
    print(f"wrote {DST}: {n} records, {len(out)} bytes raw", file=sys.stderr)


if __name__ == "__main__":
    main()
