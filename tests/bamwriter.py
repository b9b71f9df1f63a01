"""Minimal BAM (BGZF) writer for tests: enough of the format for the packers (no aux tags)."""
import struct
import zlib

_NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}


def _bgzf_block(data):
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize)
            + comp + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bam(path, refs, reads, block=40000, aux=None, align=False):
    """refs: [(name, length)]; reads: [(tid, pos0, flag, name, [(op_char, len)], seq)] in file order;
    aux: optional list of raw aux-tag bytes per read.  align: end a BGZF block rather than cut a record (what htslib
    does: the header gets its own block, a record is only split if it is larger than a block); default: blocks of
    exactly `block` bytes, records cut anywhere."""
    text = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs)))
    for name, ln in refs:
        out += struct.pack("<i", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<i", ln)
    cuts = [len(out)] if align else None                    # block boundaries (align)
    for ri, (tid, pos, flag, name, cigar, seq) in enumerate(reads):
        packed = bytearray((len(seq) + 1) // 2)
        for i, ch in enumerate(seq):
            packed[i >> 1] |= _NT16[ch] << (0 if i & 1 else 4)
        cig = b"".join(struct.pack("<I", (ln << 4) | _OPS[op]) for op, ln in cigar)
        body = (struct.pack("<iiBBHHHiiii", tid, pos, len(name) + 1, 42, 4680, len(cigar), flag, len(seq), -1, -1, 0)
                + name.encode() + b"\x00" + cig + bytes(packed) + b"\xff" * len(seq)
                + (aux[ri] if aux else b""))
        if align and len(out) + 4 + len(body) - cuts[-1] > block and len(out) > cuts[-1]:
            cuts.append(len(out))
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as fh:
        if align:
            edges = [0] + cuts + [len(out)]
            for a, b in zip(edges[:-1], edges[1:]):
                for i in range(a, b, 65280):                # (a record larger than a block is still split)
                    fh.write(_bgzf_block(bytes(out[i:min(b, i + 65280)])))
        else:
            for i in range(0, len(out), block):
                fh.write(_bgzf_block(bytes(out[i:i + block])))
        fh.write(_bgzf_block(b""))                          # BGZF EOF marker
